"""Time the VGG front-end (forward, backward) at cfg4-like shapes: N = B*T frames of [80,1,3]."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tensorflow_end2end_speech_recognition_b200 import ops


def main():
    B, T, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 1500, 80, 1
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(0)
    chans = (3, 64, 64, 128, 128)
    p = {}
    for i, n in enumerate(ops.VGG_CONVS):
        p[n + "/weight"] = torch.tensor((rng.randn(3, 3, chans[i], chans[i + 1]) * 0.05).astype(np.float32), device=dev)
        p[n + "/bias"] = torch.zeros(chans[i + 1], device=dev)
    p["bridge/weights"] = torch.tensor((rng.randn(20 * 128, 256) * 0.02).astype(np.float32), device=dev)
    p["bridge/biases"] = torch.zeros(256, device=dev)
    g = {k: torch.zeros_like(v) for k, v in p.items()}
    N = B * T
    x = torch.randn(N, H * W * 3, device=dev)
    d_out = torch.randn(N, 256, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for name, prec in (("fp32 CUDA-core", ops.PREC_FP32), ("bf16 tcgen05", ops.PREC_BF16)):
        desc = ops.vgg_desc(N, H, W, keep_prob=0.8, dropout_seed=1, precision=prec)
        for it in range(4):
            ev[0].record()
            out, reserve = ops.vgg_frontend_forward(desc, x, p)
            ev[1].record()
            ops.vgg_frontend_backward(desc, p, d_out, reserve, g)
            ev[2].record()
            torch.cuda.synchronize()
            print("%-15s N=%d fwd %.2f ms  bwd %.2f ms" % (name, N, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])),
                  flush=True)
        del out, reserve
    flops_fwd = N * (80 * (3 * 3 * 64 + 3 * 64 * 64) + 40 * (3 * 64 * 128 + 3 * 128 * 128) + 2560 * 256) * 2
    print("algorithmic fwd GFLOP %.1f (non-padding taps)" % (flops_fwd / 1e9))


main()
