"""One attention decoder step at the config-3 shape: GB/s of algorithmic bytes 4*B*T*(A+E)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.attention_layer import AttentionLayer
dev = torch.device("cuda:0")
peak = 6489.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for atype, B in (("hybrid", 64), ("bahdanau_content", 64), ("luong_general", 64), ("hybrid", 8)):
    T, E, Dq, A = 1000, 1024, 256, 128
    layer = AttentionLayer(atype, A, 0.1, 1.0, False)
    layer.create_variables(E, Dq, np.random.RandomState(0), dev)
    enc = torch.randn(B, T, E, device=dev)
    q = torch.randn(B, Dq, device=dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    pa = torch.softmax(torch.randn(B, T, device=dev), -1)
    layer.precompute_keys(enc)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        layer(enc, q, lens, pa)
    ts = []
    for _ in range(7):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); layer(enc, q, lens, pa); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    Ak = layer._keys.shape[-1] if layer._keys is not None else 0
    by = 4.0 * B * T * (Ak + E)
    print("%-18s B=%d T=%d E=%d A=%d  %.3f ms/step (incl. query GEMM)  %.0f GB/s = %.3f of measured HBM peak"
          % (atype, B, T, E, Ak, ms, by / ms / 1e6, by / ms / 1e6 / peak), flush=True)
