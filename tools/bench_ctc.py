"""CTC loss+grad at the config-2 and config-4 (CSJ-kanji, C=3001) shapes: GB/s of algorithmic bytes."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflow_end2end_speech_recognition_b200 import ops
dev = torch.device("cuda:0")
peak = 6489.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def bench(T, B, C, lmin, lmax, label):
    rng = np.random.RandomState(0)
    labels = [list(rng.randint(0, C - 1, size=int(rng.randint(lmin, lmax + 1)))) for _ in range(B)]
    flat, offs, lm = ops.pack_labels(labels)
    lg = torch.randn(T, B, C, device=dev)
    seq = torch.full((B,), T, dtype=torch.int32, device=dev)
    dflat, doffs = torch.tensor(flat, device=dev), torch.tensor(offs, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run():
        ops.ctc_loss_grad(lg, dflat, doffs, seq, lm)
    for _ in range(3):
        run()
    ts = []
    for _ in range(5):
        flush.zero_()                      # L2 flush between timed iterations
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    s_pad = (2 * lm + 1 + 31) // 32 * 32
    alg = 8.0 * T * B * C
    spill = 16.0 * T * B * s_pad
    print("%-34s T=%d B=%d C=%d Lmax=%d  %.3f ms  algorithmic 8TBC=%.1f MB -> %.0f GB/s (%.3f of measured HBM); "
          "with lattice spill %.0f GB/s; 3-pass traffic 12TBC -> %.0f GB/s" %
          (label, T, B, C, lm, ms, alg / 1e6, alg / ms / 1e6, alg / ms / 1e6 / peak,
           (alg + spill) / ms / 1e6, 12.0 * T * B * C / ms / 1e6), flush=True)


bench(1000, 64, 29, 150, 250, "cfg2 LibriSpeech char")
bench(1500, 32, 3001, 30, 70, "cfg4 CSJ kanji")
bench(300, 8, 62, 20, 40, "cfg1 TIMIT phones")
