"""Attention step kernel (b2_attention_step_forward) behind the reference-named
AttentionLayer vs oracle/attention.py for the seven implemented score types.
North-star tolerance for attention logits: 1e-3 rtol fp32 -> asserted at 2e-4."""
import numpy as np
import pytest
import torch

from oracle import attention as oatt

pytestmark = pytest.mark.gpu


def run(cuda, atype, B, T, E, Dq, A, sharpen=1.0, sigmoid=False, seed=0, prev="rand"):
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.attention_layer import AttentionLayer
    rng = np.random.RandomState(seed)
    layer = AttentionLayer(atype, A, 0.1, sharpen, sigmoid)
    layer.create_variables(E, Dq, rng, cuda)
    for k in layer.variables:                      # make biases / scales non-trivial
        if k.endswith("biases"):
            layer.variables[k] += torch.tensor(rng.randn(*layer.variables[k].shape).astype(np.float32) * 0.1, device=cuda)
    enc = rng.randn(B, T, E).astype(np.float32)
    lens = np.array([T] + [int(rng.randint(1, T + 1)) for _ in range(B - 1)], np.int32)
    for b in range(B):
        enc[b, lens[b]:] = 0
    query = rng.randn(B, Dq).astype(np.float32)
    pa = np.zeros((B, T), np.float32)
    if prev == "rand":
        pa = np.abs(rng.rand(B, T)).astype(np.float32)
        for b in range(B):
            pa[b, lens[b]:] = 0
        pa /= pa.sum(-1, keepdims=True)
    a, c = layer(torch.tensor(enc, device=cuda), torch.tensor(query, device=cuda),
                 torch.tensor(lens, device=cuda), torch.tensor(pa, device=cuda))
    torch.cuda.synchronize()
    p = {k: v.cpu().numpy() for k, v in layer.variables.items()}
    ar, cr = oatt.attention_step(atype, enc, query, lens, pa, p, sharpen, sigmoid)
    return a.cpu().numpy(), c.cpu().numpy(), ar, cr, lens


@pytest.mark.parametrize("atype", oatt.ATTENTION_TYPE)
@pytest.mark.parametrize("B,T,E,Dq,A", [(3, 50, 64, 64, 32), (5, 333, 128, 128, 48)])
def test_attention_types(cuda, atype, B, T, E, Dq, A):
    a, c, ar, cr, lens = run(cuda, atype, B, T, E, Dq, A, seed=B)
    np.testing.assert_allclose(a, ar, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(c, cr, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(a.sum(-1), 1.0, atol=1e-5)
    for b in range(B):
        assert np.all(a[b, lens[b]:] == 0)


@pytest.mark.parametrize("atype", ["hybrid", "bahdanau_content", "luong_general"])
def test_sharpening_and_sigmoid_smoothing(cuda, atype):
    a, c, ar, cr, _ = run(cuda, atype, 4, 120, 64, 64, 32, sharpen=2.0, sigmoid=False, seed=7)
    np.testing.assert_allclose(a, ar, rtol=5e-4, atol=1e-6)
    a, c, ar, cr, _ = run(cuda, atype, 4, 120, 64, 64, 32, sharpen=1.0, sigmoid=True, seed=8)
    np.testing.assert_allclose(a, ar, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(c, cr, rtol=2e-4, atol=2e-5)


def test_reference_zero_previous_weights_quirk(cuda):
    """the reference always feeds zero previous weights to location/hybrid:
    the location term then reduces to its bias -- both behaviours must match the oracle."""
    a, c, ar, cr, _ = run(cuda, "location", 3, 80, 32, 32, 16, prev="zeros", seed=9)
    np.testing.assert_allclose(a, ar, rtol=2e-4, atol=1e-6)


def test_librispeech_shape(cuda):
    """config-3 shape: B=8 per GPU, T=1000, E=1024, A=128, hybrid."""
    a, c, ar, cr, _ = run(cuda, "hybrid", 8, 1000, 1024, 256, 128, seed=11)
    np.testing.assert_allclose(a, ar, rtol=5e-4, atol=1e-6)
    np.testing.assert_allclose(c, cr, rtol=5e-4, atol=5e-5)
