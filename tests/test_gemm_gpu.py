"""b2_gemm: fp32 CUDA-core path and bf16 tcgen05 path vs fp64 numpy.

fp32 path: atol/rtol 1e-4 (summation order only).  bf16 path: operands are
rounded to bf16 by the library; the checker rounds the same way, so the only
difference left is fp32 accumulation order -> rtol 2e-3 of the row scale."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def bf16_round(a):
    return torch.tensor(a, dtype=torch.float32).to(torch.bfloat16).to(torch.float64).numpy()


def run(M, N, K, ta, tb, prec, dev, bias=True, beta=0.0, seed=0, strided=False):
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(seed)
    A = rng.randn(*((K, M) if ta else (M, K))).astype(np.float32)
    B = rng.randn(*((N, K) if tb else (K, N))).astype(np.float32)
    bvec = rng.randn(N).astype(np.float32) if bias else None
    C0 = rng.randn(M, N).astype(np.float32)
    dA, dB = torch.tensor(A, device=dev), torch.tensor(B, device=dev)
    out = torch.tensor(C0, device=dev) if beta else None
    got = ops.gemm(dA, dB, ta, tb, torch.tensor(bvec, device=dev) if bias else None, prec, out, beta)
    torch.cuda.synchronize()
    if prec == ops.PREC_BF16:
        Ar, Br = bf16_round(A), bf16_round(B)
    else:
        Ar, Br = A.astype(np.float64), B.astype(np.float64)
    ref = (Ar.T if ta else Ar) @ (Br.T if tb else Br)
    if bias:
        ref = ref + bvec
    if beta:
        ref = ref + beta * C0
    return got.cpu().numpy().astype(np.float64), ref


SHAPES = [(128, 256, 64), (200, 96, 80), (513, 300, 129), (64, 29, 1024), (1000, 2048, 592),
          (37, 61, 8)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_fp32(cuda, M, N, K, ta, tb):
    from tensorflow_end2end_speech_recognition_b200 import ops
    got, ref = run(M, N, K, ta, tb, ops.PREC_FP32, cuda)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4 * np.sqrt(K))


@pytest.mark.parametrize("M,N,K", SHAPES + [(4096, 4096, 1024), (1024, 4096, 4096)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_bf16_tcgen05(cuda, M, N, K, ta, tb):
    from tensorflow_end2end_speech_recognition_b200 import ops
    got, ref = run(M, N, K, ta, tb, ops.PREC_BF16, cuda)
    np.testing.assert_allclose(got, ref, rtol=2e-3, atol=2e-3 * np.sqrt(K))


@pytest.mark.parametrize("prec", [0, 1])
def test_beta_accumulate(cuda, prec):
    got, ref = run(300, 200, 520, 1, 0, prec, cuda, bias=False, beta=1.0)
    np.testing.assert_allclose(got, ref, rtol=2e-3, atol=2e-3 * np.sqrt(520))


def test_bf16_linearity_large(cuda):
    """size-independent property at the BASELINE GEMM shape: G(x1+x2) = G(x1)+G(x2)
    holds to fp32 rounding when x1, x2 are exactly representable in bf16."""
    from tensorflow_end2end_speech_recognition_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(0)
    M, N, K = 8192, 4096, 1024
    x1 = (torch.randint(-8, 9, (M, K), generator=g).float() / 8).to(cuda)
    x2 = (torch.randint(-8, 9, (M, K), generator=g).float() / 8).to(cuda)
    W = (torch.randint(-8, 9, (N, K), generator=g).float() / 16).to(cuda)
    y1 = ops.gemm(x1, W, False, True, None, ops.PREC_BF16)
    y2 = ops.gemm(x2, W, False, True, None, ops.PREC_BF16)
    y12 = ops.gemm(x1 + x2, W, False, True, None, ops.PREC_BF16)
    torch.cuda.synchronize()
    assert (y1 + y2 - y12).abs().max().item() < 1e-3
    # spot-check rows against fp64
    ref = x1[:4].double().cpu() @ W.double().cpu().T
    np.testing.assert_allclose(y1[:4].cpu().double().numpy(), ref.numpy(), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("M", [1, 8, 13, 32, 50, 64])
@pytest.mark.parametrize("transb", [False, True])
def test_skinny_fp32_gemm(cuda, M, transb):
    """M <= 64 takes the split-K skinny kernel (decoder-step products); any N/K, bias, beta, strides."""
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(M + 7 * transb)
    for N, K in [(1024, 1344), (30, 256), (77, 100), (1344, 1024), (64, 64)]:
        A = rng.randn(M, K + 3).astype(np.float32)[:, :K]
        Bm = rng.randn(N, K).astype(np.float32) if transb else rng.randn(K, N).astype(np.float32)
        bias = rng.randn(N).astype(np.float32)
        C0 = rng.randn(M, N).astype(np.float32)
        At = torch.tensor(np.ascontiguousarray(rng.randn(M, K + 3).astype(np.float32)), device=cuda)
        At[:, :K] = torch.tensor(A, device=cuda)
        Av = At[:, :K]                                      # strided rows
        out = torch.tensor(C0, device=cuda)
        ops.gemm(Av, torch.tensor(Bm, device=cuda), False, transb, torch.tensor(bias, device=cuda), out=out, beta=1.0)
        ref = A.astype(np.float64) @ (Bm.T if transb else Bm).astype(np.float64) + bias + C0
        np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-4, atol=1e-4 * np.sqrt(K))
        out2 = ops.gemm(Av, torch.tensor(Bm, device=cuda), False, transb)
        np.testing.assert_allclose(out2.cpu().numpy(), ref - bias - C0, rtol=1e-4, atol=1e-4 * np.sqrt(K))
