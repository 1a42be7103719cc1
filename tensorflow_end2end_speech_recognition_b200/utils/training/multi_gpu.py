"""Tower-gradient averaging, same name/signature as the reference's
``utils/training/multi_gpu.py:13-48``.

The reference concatenates every tower's gradient on ``/cpu:0`` and takes the
mean.  Here each tower is a rank (one process per GPU) or a device-local list:
``average_gradients`` keeps the list-of-towers call shape for single-process
use, ``allreduce_mean_`` is the NCCL form used by the data-parallel step
(sum over NVLink/NVSwitch, the 1/N folded into the clip kernel's post_scale).
"""
import torch


def average_gradients(total_grads_and_vars):
    """total_grads_and_vars: list (towers) of lists of (grad, var) -> list of (mean grad, var).
    Towers whose gradient is None are skipped, exactly like multi_gpu.py:30-40."""
    out = []
    for tower_grads_and_vars in zip(*total_grads_and_vars):
        grads = [g for g, _ in tower_grads_and_vars if g is not None]
        dev = grads[0].device
        mean = torch.stack([g.to(dev) for g in grads], dim=0).mean(dim=0)
        out.append((mean, tower_grads_and_vars[0][1]))
    return out


def allreduce_mean_(flat_grad, world_size, async_op=False, group=None):
    """In-place all-reduce(sum) of an (already 1/N-scaled) flat fp32 gradient bucket."""
    import torch.distributed as dist
    if world_size <= 1:
        return None
    return dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
