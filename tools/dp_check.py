"""Data-parallel step on N GPUs of one box (torchrun): NCCL through the C ABI (b2_allreduce_mean), per-layer
buckets overlapped with BPTT vs one all-reduce after it vs torch.distributed -- same parameters after a step,
and step times.  Usage: torchrun --nproc-per-node N tools/dp_check.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC                      # noqa: E402
from tensorflow_end2end_speech_recognition_b200.utils.training.multi_gpu import NcclComm       # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = NcclComm(rank, world, device=dev)
    B, T, D = (int(os.environ.get("DP_B", "64")), int(os.environ.get("DP_T", "1000")), 80)
    rng = np.random.RandomState(100 + rank)
    x = rng.randn(B, T, D).astype(np.float32)
    seq = np.full(B, T, np.int32)
    labels = [list(rng.randint(0, 28, size=int(rng.randint(150, 250) * T / 1000) + 1)) for _ in range(B)]
    results = {}
    for mode in ("torch", "c_abi_single", "c_abi_buckets"):
        model = CTC(encoder_type="blstm", input_size=D, num_units=512, num_layers=5, num_classes=28,
                    clip_grad_norm=5.0, precision="bf16", device=dev, seed=1)
        model.set_data_parallel(world, comm=None if mode == "torch" else comm)
        model.bucketed_exchange = (mode == "c_abi_buckets")
        xd, sd = torch.tensor(x, device=dev), torch.tensor(seq, device=dev)

        def step():
            loss, _ = model.compute_loss(xd, labels, sd, keep_prob=1.0)
            model.train(loss, "sgd", 1e-3)
        step()
        torch.cuda.synchronize()
        p1 = model.flat_params.clone()
        for _ in range(3):
            step()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            step()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 8], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # every rank must hold identical parameters
        ref = p1.clone()
        dist.broadcast(ref, src=0)
        same = bool((ref == p1).all())
        results[mode] = (float(ms), p1, same)
        names = [(v.name, (v.tensor.data_ptr() - model.flat_params.data_ptr()) // 4, v.tensor.numel())
                 for v in model.trainable_variables()]
        if rank == 0:
            print("%-14s %.3f ms/step  ranks identical after step: %s" % (mode, float(ms), same), flush=True)
    if rank == 0:
        a, b, c = results["torch"][1], results["c_abi_single"][1], results["c_abi_buckets"][1]
        d1 = float((a - b).abs().max() / a.abs().max())
        d2 = float((a - c).abs().max() / a.abs().max())
        print("max |param diff| / max|param| after one step: torch vs c_abi_single %.2e, torch vs buckets %.2e" % (d1, d2))
        if d2 >= 1e-5:
            for n, o, k in names:
                print("   %-50s %.3e" % (n, float((a[o:o + k] - c[o:o + k]).abs().max())))
        assert d1 < 1e-5 and d2 < 1e-5, (d1, d2)
        print("dp_check ok")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
