"""BASELINE config 5: greedy + beam-20 CTC decode of 256 utterances (T=1000, C=29)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflow_end2end_speech_recognition_b200 import ops
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
Bn, T, C = 256, 1000, 29
x = rng.randn(Bn, T, C).astype(np.float32) * 4.0
x[..., C - 1] += 3.0                       # blank-heavy, peaky posteriors like a trained model
lp = torch.log_softmax(torch.tensor(x), -1)
seq = torch.full((Bn,), T, dtype=torch.int32, device=dev)
lp_d = lp.to(dev)
logits_tbc = lp_d.transpose(0, 1).contiguous()


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


ms_g = timeit(lambda: ops.ctc_greedy_decode(logits_tbc, seq))
ms_b = timeit(lambda: ops.ctc_beam_decode(lp_d, seq, 20))
print("greedy: %.3f ms for %d utt -> %.0f utt/s ; beam 20: %.1f ms -> %.0f utt/s" %
      (ms_g, Bn, Bn / ms_g * 1e3, ms_b, Bn / ms_b * 1e3), flush=True)
# (label parity of these decoders vs the oracle / the reference's golden vectors: tests/test_decode_gpu.py)
