"""One CTC loss+grad call at a BASELINE shape (for ncu): prof_ctc_one.py cfg2|cfg4"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflow_end2end_speech_recognition_b200 import ops
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
T, B, C, lmin, lmax = (1000, 64, 29, 150, 250) if which == "cfg2" else (1500, 32, 3001, 30, 70)
rng = np.random.RandomState(0)
labels = [list(rng.randint(0, C - 1, size=int(rng.randint(lmin, lmax + 1)))) for _ in range(B)]
flat, offs, lm = ops.pack_labels(labels)
lg = torch.randn(T, B, C, device=dev)
seq = torch.full((B,), T, dtype=torch.int32, device=dev)
dflat, doffs = torch.tensor(flat, device=dev), torch.tensor(offs, device=dev)
for _ in range(2):
    ops.ctc_loss_grad(lg, dflat, doffs, seq, lm)
torch.cuda.synchronize()
