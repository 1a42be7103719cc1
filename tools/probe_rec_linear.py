"""Forward / forward+backward time of one 512-unit layer (tiny input width: the gate GEMM is negligible) against T:
fixed cost vs per-step cost of the recurrence kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflow_end2end_speech_recognition_b200 import ops
dev = torch.device("cuda:0")


def run(T, B, backward):
    D, H = 64, 512
    rng = np.random.RandomState(0)
    P, G = {}, {}
    for d in ("fw", "bw"):
        P[d] = {"kernel": torch.tensor(rng.uniform(-0.1, 0.1, (D + H, 4 * H)).astype(np.float32), device=dev),
                "bias": torch.zeros(4 * H, device=dev)}
        for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
            P[d][k] = torch.tensor(rng.uniform(-0.1, 0.1, H).astype(np.float32), device=dev)
        G[d] = {k: torch.zeros_like(v) for k, v in P[d].items()}
    x = torch.randn(T, B, D, device=dev)
    dy = torch.randn(T, B, 2 * H, device=dev)
    seq = torch.full((B,), T, dtype=torch.int32, device=dev)
    desc = ops.lstm_desc(T, B, D, H, precision=ops.PREC_BF16, need_backward=backward)

    def go():
        y, fs, res = ops.blstm_layer_forward(desc, x, seq, P["fw"], P["bw"])
        if backward:
            ops.blstm_layer_backward(desc, x, seq, P["fw"], P["bw"], dy, res, G["fw"], G["bw"])
            ops.blstm_backward_join()
    for _ in range(2):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        go()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


for B in (64, 32):
    for bwd in (False, True):
        ts = [(T, run(T, B, bwd)) for T in (125, 250, 500, 1000, 2000)]
        a = (ts[-1][1] - ts[1][1]) / (ts[-1][0] - ts[1][0])
        b = ts[-1][1] - a * ts[-1][0]
        print("B=%d %s: %s  -> %.3f us/step + %.3f ms fixed" % (B, "fwd+bwd" if bwd else "fwd    ",
              "  ".join("T=%d %.3f ms" % t for t in ts), a * 1e3, b), flush=True)
