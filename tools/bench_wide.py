"""Times one wide BLSTM layer (config 4: H = 1024, T = 1500, B = 32) forward / forward+backward, with the grid-resident
recurrence (lstm_wide.cu) and with the per-frame fallback."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflow_end2end_speech_recognition_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def bench(T, B, D, H, backward, label, iters=3):
    rng = np.random.RandomState(0)
    P, G = {}, {}
    for d in ("fw", "bw"):
        P[d] = {"kernel": torch.tensor(rng.uniform(-0.05, 0.05, (D + H, 4 * H)).astype(np.float32), device=dev),
                "bias": torch.zeros(4 * H, device=dev)}
        for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
            P[d][k] = torch.tensor(rng.uniform(-0.05, 0.05, H).astype(np.float32), device=dev)
        G[d] = {k: torch.zeros_like(v) for k, v in P[d].items()}
    x = torch.randn(T, B, D, device=dev)
    dy = torch.randn(T, B, 2 * H, device=dev)
    seq = torch.full((B,), T, dtype=torch.int32, device=dev)
    desc = ops.lstm_desc(T, B, D, H, precision=ops.PREC_BF16, need_backward=backward)

    def run():
        y, fs, res = ops.blstm_layer_forward(desc, x, seq, P["fw"], P["bw"])
        if backward:
            ops.blstm_layer_backward(desc, x, seq, P["fw"], P["bw"], dy, res, G["fw"], G["bw"])
            ops.blstm_backward_join()
    for _ in range(4):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%-46s T=%d B=%d D=%d H=%d  %.2f ms/layer (%.2f us/frame)" % (label, T, B, D, H, ms, ms * 1e3 / T), flush=True)


if __name__ == "__main__":
    T, B = int(os.environ.get("WIDE_T", "1500")), int(os.environ.get("WIDE_B", "32"))
    for wide in (("1",) if os.environ.get("B2_WIDE_ONLY") else ("1", "0")):
        os.environ["B2_WIDE_REC"] = wide
        tag = "grid-resident" if wide == "1" else "per-frame fallback"
        bench(T, B, 2048, 1024, False, "fwd only, " + tag)
        bench(T, B, 2048, 1024, True, "fwd+bwd, " + tag)
