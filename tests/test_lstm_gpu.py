"""b2_blstm_layer_forward/backward vs the torch-CPU oracle (oracle/lstm.py).

fp32 path: encoder states within rtol 1e-3 (north star) -- asserted at 1e-4 --
and gradients within 1e-3 of the autograd oracle.  bf16 (tcgen05) path: same
checks with the tolerance a bf16-operand GEMM allows (documented in DESIGN.md)."""
import numpy as np
import pytest
import torch

from oracle import lstm as olstm

pytestmark = pytest.mark.gpu


def to_dev(layer, dev):
    return {d: {k: torch.tensor(v, dtype=torch.float32, device=dev) for k, v in layer[d].items()}
            for d in ("fw", "bw")}


def oracle_layer(x_tbd, seq, layer, dy, cell_clip=None, masks=None, keep_prob=1.0):
    p = {d: {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in layer[d].items()}
         for d in ("fw", "bw")}
    x = torch.tensor(x_tbd.astype(np.float64), requires_grad=True)
    y, fs = olstm.blstm_forward(x.transpose(0, 1), seq, [p], keep_prob=keep_prob,
                                dropout_masks=masks, cell_clip=cell_clip)
    (y * torch.tensor(dy.astype(np.float64))).sum().backward()
    grads = {d: {k: v.grad.numpy() for k, v in p[d].items()} for d in p}
    return y.detach().numpy(), fs, x.grad.numpy(), grads


def run_layer(dev, T, B, D, H, seq, precision, peephole=True, cell_clip=None, seed=0, keep_prob=1.0,
              parameter_init=0.3):
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(seed)
    layer = olstm.init_blstm_params(D, H, 1, parameter_init=parameter_init, use_peephole=peephole, seed=seed)[0]
    for d in layer:
        layer[d]["bias"] = (rng.randn(4 * H) * 0.1).astype(np.float32)
    x = rng.randn(T, B, D).astype(np.float32)
    dy = rng.randn(T, B, 2 * H).astype(np.float32)
    P = to_dev(layer, dev)
    desc = ops.lstm_desc(T, B, D, H, use_peephole=peephole, cell_clip=cell_clip, precision=precision,
                         keep_prob=keep_prob, dropout_seed=1234)
    seq_t = torch.tensor(np.asarray(seq, np.int32), device=dev)
    xd = torch.tensor(x, device=dev)
    y, fs, reserve = ops.blstm_layer_forward(desc, xd, seq_t, P["fw"], P["bw"], want_final_state=True)
    G = {d: {k: torch.zeros_like(v) for k, v in P[d].items()} for d in P}
    dx = ops.blstm_layer_backward(desc, xd, seq_t, P["fw"], P["bw"], torch.tensor(dy, device=dev),
                                  reserve, G["fw"], G["bw"])
    ops.blstm_backward_join()
    torch.cuda.synchronize()
    masks = None
    if keep_prob < 1.0:
        from tests.util_dropout import dropout_mask
        masks = [torch.tensor(dropout_mask(1234, T * B * 2 * H, keep_prob).reshape(T, B, 2 * H))]
    yr, fsr, dxr, gr = oracle_layer(x, seq, layer, dy, cell_clip, masks, keep_prob)
    return (y.cpu().numpy(), fs.cpu().numpy(), dx.cpu().numpy(),
            {d: {k: v.cpu().numpy() for k, v in G[d].items()} for d in G}), (yr, fsr, dxr, gr)


def compare(got, ref, tol_y, tol_g):
    y, fs, dx, g = got
    yr, fsr, dxr, gr = ref
    np.testing.assert_allclose(y, yr, rtol=tol_y, atol=tol_y)
    np.testing.assert_allclose(fs[0], fsr[0][0].detach().numpy(), rtol=tol_y, atol=tol_y)
    np.testing.assert_allclose(fs[1], fsr[0][1].detach().numpy(), rtol=tol_y, atol=tol_y)
    np.testing.assert_allclose(fs[2], fsr[1][0].detach().numpy(), rtol=tol_y, atol=tol_y)
    np.testing.assert_allclose(fs[3], fsr[1][1].detach().numpy(), rtol=tol_y, atol=tol_y)
    scale = max(1.0, np.abs(dxr).max())
    np.testing.assert_allclose(dx, dxr, rtol=tol_g, atol=tol_g * scale)
    for d in g:
        for k in g[d]:
            s = max(1.0, np.abs(gr[d][k]).max())
            np.testing.assert_allclose(g[d][k], gr[d][k], rtol=tol_g, atol=tol_g * s, err_msg="%s/%s" % (d, k))


@pytest.mark.parametrize("T,B,D,H,seq", [
    (7, 3, 6, 8, [7, 4, 5]),
    (20, 5, 24, 32, [20, 20, 13, 7, 1]),
    (33, 17, 40, 48, None),
])
@pytest.mark.parametrize("peephole", [True, False])
def test_layer_fp32(cuda, T, B, D, H, seq, peephole):
    from tensorflow_end2end_speech_recognition_b200 import ops
    if seq is None:
        seq = [T] + list(np.random.RandomState(1).randint(1, T + 1, size=B - 1))
    got, ref = run_layer(cuda, T, B, D, H, seq, ops.PREC_FP32, peephole=peephole)
    compare(got, ref, 1e-4, 1e-3)


def test_layer_fp32_cell_clip(cuda):
    from tensorflow_end2end_speech_recognition_b200 import ops
    got, ref = run_layer(cuda, 25, 4, 16, 16, [25, 20, 11, 3], ops.PREC_FP32, cell_clip=0.5, seed=3)
    compare(got, ref, 1e-4, 1e-3)


def test_layer_fp32_dropout(cuda):
    from tensorflow_end2end_speech_recognition_b200 import ops
    got, ref = run_layer(cuda, 12, 4, 16, 16, [12, 9, 12, 5], ops.PREC_FP32, keep_prob=0.8, seed=5)
    compare(got, ref, 1e-4, 1e-3)


@pytest.mark.parametrize("T,B,D,H", [(24, 16, 80, 64), (40, 32, 128, 128)])
def test_layer_bf16(cuda, T, B, D, H):
    from tensorflow_end2end_speech_recognition_b200 import ops
    seq = [T] + list(np.random.RandomState(2).randint(T // 2, T + 1, size=B - 1))
    got, ref = run_layer(cuda, T, B, D, H, seq, ops.PREC_BF16, seed=7)
    compare(got, ref, 3e-2, 5e-2)


@pytest.mark.parametrize("T,B,D,H,P,seq", [
    (6, 3, 5, 8, 4, [6, 3, 5]),
    (15, 7, 20, 32, 12, None),
    (9, 70, 16, 24, 24, None),          # batch > 64: the projection products fall back to the tile GEMM
])
@pytest.mark.parametrize("keep_prob", [1.0, 0.7])
def test_layer_num_proj(cuda, T, B, D, H, P, seq, keep_prob):
    """LSTMCell(num_proj) (blstm.py:215-228, lstm.py:171-176): projected recurrent / emitted state,
    forward + BPTT incl. d(projection), dropout applied after the projection."""
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(T * 100 + B)
    if seq is None:
        seq = [T] + list(rng.randint(1, T + 1, size=B - 1))
    layer = olstm.init_blstm_params(D, H, 1, parameter_init=0.3, use_peephole=True, num_proj=P, seed=3)[0]
    for d in layer:
        layer[d]["bias"] = (rng.randn(4 * H) * 0.1).astype(np.float32)
    x = rng.randn(T, B, D).astype(np.float32)
    dy = rng.randn(T, B, 2 * P).astype(np.float32)
    Pd = to_dev(layer, cuda)
    desc = ops.lstm_desc(T, B, D, H, use_peephole=True, cell_clip=3.0, precision=ops.PREC_FP32, keep_prob=keep_prob,
                         dropout_seed=77, num_proj=P)
    seq_t = torch.tensor(np.asarray(seq, np.int32), device=cuda)
    xd = torch.tensor(x, device=cuda)
    y, fs, reserve = ops.blstm_layer_forward(desc, xd, seq_t, Pd["fw"], Pd["bw"], want_final_state=True)
    G = {d: {k: torch.zeros_like(v) for k, v in Pd[d].items()} for d in Pd}
    dx = ops.blstm_layer_backward(desc, xd, seq_t, Pd["fw"], Pd["bw"], torch.tensor(dy, device=cuda), reserve,
                                  G["fw"], G["bw"])
    torch.cuda.synchronize()
    masks = None
    if keep_prob < 1.0:
        from tests.util_dropout import dropout_mask
        masks = [torch.tensor(dropout_mask(77, T * B * 2 * P, keep_prob).reshape(T, B, 2 * P))]
    yr, fsr, dxr, gr = oracle_layer(x, seq, layer, dy, 3.0, masks, keep_prob)
    assert y.shape == (T, B, 2 * P)
    np.testing.assert_allclose(y.cpu().numpy(), yr, rtol=1e-4, atol=1e-4)
    for got, ref in zip(fs, (fsr[0][0], fsr[0][1], fsr[1][0], fsr[1][1])):
        np.testing.assert_allclose(got.cpu().numpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dx.cpu().numpy(), dxr, rtol=1e-3, atol=1e-3 * max(1.0, np.abs(dxr).max()))
    for d in G:
        assert set(G[d]) == set(gr[d])
        for k in G[d]:
            s = max(1.0, np.abs(gr[d][k]).max())
            np.testing.assert_allclose(G[d][k].cpu().numpy(), gr[d][k], rtol=1e-3, atol=1e-3 * s, err_msg="%s/%s" % (d, k))
