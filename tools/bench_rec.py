"""Times one BLSTM layer (forward, optionally backward) at the BASELINE config-2 shape."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflow_end2end_speech_recognition_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def bench(T, B, D, H, prec, backward, iters=5, label="", keep_prob=1.0):
    rng = np.random.RandomState(0)
    P = {}
    for d in ("fw", "bw"):
        P[d] = {"kernel": torch.tensor(rng.uniform(-0.1, 0.1, (D + H, 4 * H)).astype(np.float32), device=dev),
                "bias": torch.zeros(4 * H, device=dev)}
        for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
            P[d][k] = torch.tensor(rng.uniform(-0.1, 0.1, H).astype(np.float32), device=dev)
    G = {d: {k: torch.zeros_like(v) for k, v in P[d].items()} for d in P}
    x = torch.randn(T, B, D, device=dev)
    dy = torch.randn(T, B, 2 * H, device=dev)
    seq = torch.full((B,), T, dtype=torch.int32, device=dev)
    desc = ops.lstm_desc(T, B, D, H, precision=prec, need_backward=backward, keep_prob=keep_prob, dropout_seed=7)

    def run():
        y, fs, res = ops.blstm_layer_forward(desc, x, seq, P["fw"], P["bw"])
        if backward:
            ops.blstm_layer_backward(desc, x, seq, P["fw"], P["bw"], dy, res, G["fw"], G["bw"])
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%-40s T=%d B=%d D=%d H=%d  %.3f ms/layer  (%.3f us/step)" %
          (label, T, B, D, H, ms, ms * 1e3 / T), flush=True)


if __name__ == "__main__":
    bwd = "--bwd" in sys.argv
    if "--quick" in sys.argv:
        tag = " ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("B2_REC"))
        bench(1000, 64, 1024, 512, ops.PREC_BF16, True, label="fwd+bwd " + tag)
        bench(1000, 64, 1024, 512, ops.PREC_BF16, False, label="fwd only " + tag)
        sys.exit(0)
    for nch in (1, 2):
        os.environ["B2_REC_NCHAIN"] = str(nch)
        bench(1000, 64, 1024, 512, ops.PREC_BF16, bwd, label="bf16 tc nchain=%d %s" % (nch, "fwd+bwd" if bwd else "fwd"))
    os.environ.pop("B2_REC_NCHAIN")
    bench(1000, 64, 1024, 512, ops.PREC_BF16, bwd, label="bf16 tc keep_prob=0.8", keep_prob=0.8)
    bench(1000, 64, 1024, 512, ops.PREC_BF16, False, label="bf16 tc keep_prob=0.8 fwd only", keep_prob=0.8)
    bench(1000, 64, 1024, 512, ops.PREC_BF16, False, label="bf16 tc keep_prob=1.0 fwd only")
    bench(1000, 16, 1024, 512, ops.PREC_BF16, bwd, label="bf16 tc B=16")
    bench(1000, 32, 1024, 512, ops.PREC_BF16, bwd, label="bf16 tc B=32")
    bench(300, 8, 120, 256, ops.PREC_BF16, bwd, label="bf16 tc cfg1")
