"""Drivers for ncu captures of the kernels outside the headline step: which = attn | gemv | inputs | vgg | seqloss"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tensorflow_end2end_speech_recognition_b200 import ops, _lib
dev = torch.device("cuda:0")
which = sys.argv[1]
rng = np.random.RandomState(0)
lib = _lib.load()
if which == "attn":
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.attention_layer import AttentionLayer
    B, T, E, A, Hd = 64, 1000, 1024, 128, 256
    layer = AttentionLayer("hybrid", A, 0.1, 1.0, False)
    layer.create_variables(E, Hd, rng, dev)
    enc = torch.randn(B, T, E, device=dev)
    enc_len = torch.full((B,), T, dtype=torch.int32, device=dev)
    layer.precompute_keys(enc)
    h = torch.randn(B, Hd, device=dev)
    outd = {"q": torch.empty(B, A, device=dev)}
    for _ in range(2):
        alpha, ctx = layer(enc, h, enc_len, torch.zeros(B, T, device=dev), out=outd)
    g = {k: torch.zeros_like(v) for k, v in layer.variables.items()}
    d_keys = torch.zeros(B, T, A, device=dev)
    dq = torch.empty(B, A, device=dev)
    for _ in range(2):
        layer.backward_step(enc, outd["q"], alpha, None, enc_len, torch.randn(B, E, device=dev), d_keys, dq, g)
elif which == "gemv":
    for M in (8, 64):
        a, bm = torch.randn(M, 1344, device=dev), torch.randn(1344, 1024, device=dev)
        for _ in range(2):
            ops.gemm(a, bm)
elif which == "inputs":
    x2 = torch.randn(32, 1500, 240, device=dev)
    l2 = torch.full((32,), 1500, dtype=torch.int32, device=dev)
    Dout = lib.b2_stack_splice_out_dim(240, 1, 11)
    out2 = torch.empty(32, 1500, Dout, device=dev)
    ol2 = torch.empty(32, dtype=torch.int32, device=dev)
    for _ in range(2):
        lib.b2_stack_splice(ops._ptr(x2), ops._ptr(l2), 32, 1500, 240, 1, 1, 11, 1500, ops._ptr(out2), ops._ptr(ol2), ops._stream())
elif which == "seqloss":
    logits = torch.randn(64, 200, 3002, device=dev)
    tg = torch.randint(0, 3002, (64, 201), dtype=torch.int32, device=dev)
    lens = torch.full((64,), 200, dtype=torch.int32, device=dev)
    for _ in range(2):
        ops.sequence_loss(logits, tg[:, 1:], lens)
torch.cuda.synchronize()
print("done", which)
