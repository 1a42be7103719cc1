"""Tower-gradient averaging, same name/signature as the reference's
``utils/training/multi_gpu.py:13-48``.

The reference concatenates every tower's gradient on ``/cpu:0`` and takes the mean.  Here:

* ``average_gradients(total_grads_and_vars)`` keeps the list-of-towers call shape of
  ``examples/librispeech/training/train_ctc.py:143``.  Towers produced by
  ``Optimizer.compute_gradients`` carry their flat gradient buffer, so the mean of N towers is ONE
  streaming launch of ``b2_tower_mean`` over the whole 110 MB parameter set (not a per-variable
  ``concat`` + ``reduce_mean``); arbitrary ``(grad, var)`` lists take the same kernel per variable.
  In graph mode (towers built from placeholders) it returns a lazy handle for ``apply_gradients``.
* ``NcclComm`` / ``allreduce_mean_`` are the one-rank-per-GPU form used by the data-parallel step:
  ``b2_allreduce_mean`` (NCCL ``ncclAvg`` over NVLink/NVSwitch, bound from C) on per-layer buckets.
"""
import ctypes as C

import torch

from ... import _lib
from ... import ops
from ...compat import graph as _graph


def _mean_towers(towers):
    first = towers[0]
    flats = [getattr(t, "flat", None) for t in towers]
    if all(f is not None for f in flats):
        # every tower is a full flat gradient buffer with the same layout: one launch
        if len(towers) > 1:
            ops.tower_mean(flats, flats[0])
        return first
    out = []
    for tower_grads_and_vars in zip(*towers):
        grads = [g for g, _ in tower_grads_and_vars if g is not None]   # multi_gpu.py:30-40: None towers are skipped
        var = tower_grads_and_vars[0][1]
        if not grads:
            out.append((None, var))
            continue
        dev = grads[0].device
        srcs = [g.to(dev).contiguous() for g in grads]
        dst = torch.empty_like(srcs[0])
        ops.tower_mean(srcs, dst)
        out.append((dst, var))
    return out


def average_gradients(total_grads_and_vars):
    """total_grads_and_vars: list (towers) of lists of (grad, var) -> list of (mean grad, var)."""
    if any(isinstance(t, _graph.Tensor) for t in total_grads_and_vars):
        lazy = _graph.LazyGradsAndVars(
            _graph.Op(lambda towers: _mean_towers(towers), ([t.op if isinstance(t, _graph.LazyGradsAndVars) else t
                                                           for t in total_grads_and_vars],), {},
                      name="average_gradients"))
        return lazy
    return _mean_towers(list(total_grads_and_vars))


class NcclComm(object):
    """One NCCL communicator rank for this process' GPU, created through the C ABI
    (``b2_comm_get_unique_id`` on rank 0 -> id shipped over ``torch.distributed`` -> ``b2_comm_init_rank``)."""

    def __init__(self, rank, world_size, group=None, device=None):
        import torch.distributed as dist
        lib = _lib.load()
        if not lib.b2_comm_available():
            raise RuntimeError("NCCL is not available to libb2asr.so (libnccl.so.2 not found)")
        self.rank, self.world_size = int(rank), int(world_size)
        idbuf = (C.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(lib.b2_comm_get_unique_id(idbuf), "b2_comm_get_unique_id")
        backend = dist.get_backend(group)
        t = torch.tensor(list(idbuf), dtype=torch.uint8)
        if backend == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.broadcast(t, src=0, group=group)
        idbytes = bytes(t.cpu().tolist())
        self._h = C.c_void_p()
        _lib.check(lib.b2_comm_init_rank(C.byref(self._h), self.world_size, idbytes, self.rank), "b2_comm_init_rank")

    def allreduce_mean_(self, buckets, stream=None):
        """in-place mean over ranks of a list of contiguous fp32 cuda tensors (one NCCL group call)"""
        lib = _lib.load()
        n = len(buckets)
        ptrs = (C.c_void_p * n)(*[b.data_ptr() for b in buckets])
        sizes = (C.c_int64 * n)(*[b.numel() for b in buckets])
        s = C.c_void_p(stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.b2_allreduce_mean(self._h, ptrs, sizes, n, s), "b2_allreduce_mean")

    def close(self):
        if self._h:
            _lib.load().b2_comm_destroy(self._h)
            self._h = C.c_void_p()


def allreduce_mean_(flat_grad, world_size, async_op=False, group=None):
    """``torch.distributed`` form (gloo on CPU boxes, or when no NcclComm was attached): in-place
    all-reduce(sum) of an (already 1/N-scaled) flat fp32 gradient bucket."""
    import torch.distributed as dist
    if world_size <= 1:
        return None
    return dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
