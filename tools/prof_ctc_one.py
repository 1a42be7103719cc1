import sys, numpy as np, torch
sys.path.insert(0, ".")
from tensorflow_end2end_speech_recognition_b200 import ops
dev = torch.device("cuda:0")
T, B, C = 1000, 64, 29
rng = np.random.RandomState(0)
labels = [list(rng.randint(0, C - 1, size=int(rng.randint(150, 251)))) for _ in range(B)]
flat, offs, lm = ops.pack_labels(labels)
lg = torch.randn(T, B, C, device=dev)
seq = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(2):
    ops.ctc_loss_grad(lg, torch.tensor(flat, device=dev), torch.tensor(offs, device=dev), seq, lm)
torch.cuda.synchronize()
