"""Times the tcgen05 GEMM (b2_gemm_bf16) on the headline BLSTM shapes; prints TFLOP/s
and fraction of the measured bf16 peak (MEASURED_PEAKS.json)."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflow_end2end_speech_recognition_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
peak = 1653.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass


def bench(name, a_mn, b_mn, M, N, K, out_mode, ksplit=0, iters=10):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    Cm = torch.zeros((M, N), device=dev, dtype=torch.bfloat16 if out_mode == 2 else torch.float32)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        rc = lib.b2_gemm_bf16(a_mn, b_mn, M, N, K, 1.0, C.c_void_p(A.data_ptr()), A.stride(0),
                              C.c_void_p(B.data_ptr()), B.stride(0), C.c_void_p(Cm.data_ptr()), N,
                              C.c_void_p(0), out_mode, ksplit, st)
        _lib.check(rc, name)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    print("%-28s M=%6d N=%5d K=%6d  %8.3f ms  %7.1f TFLOP/s  %.3f of measured peak" %
          (name, M, N, K, ms, tf, tf / peak), flush=True)


if __name__ == "__main__":
    TB = 64000
    bench("fwd inproj L2-5 (TN)", 0, 0, TB, 4096, 1024, 0)
    bench("fwd inproj L2-5 bf16 out", 0, 0, TB, 4096, 1024, 2)
    bench("fwd inproj L1 (TN)", 0, 0, TB, 4096, 80, 0)
    bench("dX (TN)", 0, 0, TB, 1024, 4096, 0)
    bench("wgrad dWx (NT, atomic)", 1, 1, 1024, 4096, TB, 1)
    bench("wgrad dWh (NT, atomic)", 1, 1, 512, 2048, TB, 1)
    bench("square 8192", 0, 0, 8192, 8192, 8192, 0)
    bench("fc out (N=29)", 0, 1, TB, 29 + 3, 1024, 0)
