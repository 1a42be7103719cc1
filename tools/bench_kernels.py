"""CUDA-event timings of the kernels outside the headline step (attention backward, decoder-step GEMMs,
sequence loss, input pipeline, edit distance, VGG front-end) at representative sizes, against their
algorithmic bytes / FLOPs.  L2 is flushed before every timed launch.  Output -> profiles/r01_kernels_extra.log"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tensorflow_end2end_speech_recognition_b200 import ops, _lib
from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.attention_layer import AttentionLayer

dev = torch.device("cuda:0")
HBM = 6489.9
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timed(fn, n=5):
    ts = []
    for _ in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


def line(name, ms, nbytes=None, flops=None):
    s = "%-58s %9.3f ms" % (name, ms)
    if nbytes:
        gbs = nbytes / ms / 1e6
        s += "  %8.1f MB  %7.1f GB/s  %5.1f %% of HBM peak" % (nbytes / 1e6, gbs, 100 * gbs / HBM)
    if flops:
        s += "  %7.2f TFLOP/s" % (flops / ms / 1e9)
    print(s)


rng = np.random.RandomState(0)
# ---- input pipeline: 64 utterances of 3000 raw frames x 80, stack 3 / skip 3 -> [64,1000,240]
from tensorflow_end2end_speech_recognition_b200.utils.io.inputs.pipeline import DeviceInputPipeline
lib = _lib.load()
B, Traw, D, S, K = 64, 3000, 80, 3, 3
raw = torch.randn(B, Traw, D, device=dev)
raw_len = torch.full((B,), Traw, dtype=torch.int32, device=dev)
Tout, Dout = 1000, lib.b2_stack_splice_out_dim(D, S, 1)
out = torch.empty(B, Tout, Dout, device=dev)
out_len = torch.empty(B, dtype=torch.int32, device=dev)
f = lambda: lib.b2_stack_splice(ops._ptr(raw), ops._ptr(raw_len), B, Traw, D, S, K, 1, Tout, ops._ptr(out), ops._ptr(out_len), ops._stream())
line("b2_stack_splice 64x3000x80 stack3/skip3", timed(f), nbytes=raw.numel() * 4 + out.numel() * 4)
Dout2 = lib.b2_stack_splice_out_dim(240, 1, 11)
x2 = torch.randn(32, 1500, 240, device=dev)
l2 = torch.full((32,), 1500, dtype=torch.int32, device=dev)
out2 = torch.empty(32, 1500, Dout2, device=dev)
ol2 = torch.empty(32, dtype=torch.int32, device=dev)
f = lambda: lib.b2_stack_splice(ops._ptr(x2), ops._ptr(l2), 32, 1500, 240, 1, 1, 11, 1500, ops._ptr(out2), ops._ptr(ol2), ops._stream())
line("b2_stack_splice 32x1500x240 splice 11", timed(f), nbytes=x2.numel() * 4 + out2.numel() * 4)

# ---- attention backward step, cfg3 shapes at B=64 and B=8
for Bq in (64, 8):
    T, E, A, Hd = 1000, 1024, 128, 256
    layer = AttentionLayer("hybrid", A, 0.1, 1.0, False)
    layer.create_variables(E, Hd, rng, dev)
    enc = torch.randn(Bq, T, E, device=dev)
    enc_len = torch.full((Bq,), T, dtype=torch.int32, device=dev)
    layer.precompute_keys(enc)
    h = torch.randn(Bq, Hd, device=dev)
    outd = {"q": torch.empty(Bq, A, device=dev)}
    alpha, ctx = layer(enc, h, enc_len, torch.zeros(Bq, T, device=dev), out=outd)
    line("attention step forward (hybrid, zero prev) B=%d" % Bq,
         timed(lambda: layer(enc, h, enc_len, torch.zeros(Bq, T, device=dev), out=outd)),
         nbytes=4 * Bq * T * (A + E))
    dctx = torch.randn(Bq, E, device=dev)
    d_keys = torch.zeros(Bq, T, A, device=dev)
    dq = torch.empty(Bq, A, device=dev)
    g = {k: torch.zeros_like(v) for k, v in layer.variables.items()}
    line("attention step backward (hybrid) B=%d" % Bq,
         timed(lambda: layer.backward_step(enc, outd["q"], alpha, None, enc_len, dctx, d_keys, dq, g)),
         nbytes=4 * Bq * T * (E + 3 * A))

# ---- decoder-step GEMMs (skinny path)
for M in (8, 32, 64):
    Kd, N = 1344, 1024
    a = torch.randn(M, Kd, device=dev)
    bm = torch.randn(Kd, N, device=dev)
    line("cell pre-activation GEMM [%dx1344].[1344x1024] fp32" % M, timed(lambda: ops.gemm(a, bm)),
         nbytes=4 * (Kd * N + M * Kd + M * N), flops=2.0 * M * Kd * N)

# ---- sequence loss
Bq, L, C = 64, 200, 3002
logits = torch.randn(Bq, L, C, device=dev)
tg = torch.randint(0, C, (Bq, L + 1), dtype=torch.int32, device=dev)
lens = torch.full((Bq,), L, dtype=torch.int32, device=dev)
line("b2_sequence_loss B=64 L=200 C=3002 (loss + grad)", timed(lambda: ops.sequence_loss(logits, tg[:, 1:], lens)),
     nbytes=8 * Bq * L * C)

# ---- edit distance: 256 pairs of ~200 labels
hyp = [list(rng.randint(0, 28, 200)) for _ in range(256)]
ref = [list(rng.randint(0, 28, 200)) for _ in range(256)]
line("b2_edit_distance 256 pairs 200x200 (incl. H2D/D2H)", timed(lambda: ops.edit_distance(hyp, ref, dev)))
