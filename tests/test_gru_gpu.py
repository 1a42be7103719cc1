"""GRU layers (csrc/gru.cu: b2_bgru_layer_forward/backward) and the 'bgru' / 'gru' encoders (reference
models/encoders/core/gru.py): layer outputs, final state and every gradient vs the fp64 oracle's GRUCell
(oracle/lstm.py::gru_cell_step), ragged lengths; CTC model with both encoder types."""
import numpy as np
import pytest
import torch

from oracle import lstm as olstm
from oracle import model as omodel

pytestmark = pytest.mark.gpu


def _params(rng, D, H, a=0.3):
    return {"gates/kernel": rng.uniform(-a, a, (D + H, 2 * H)).astype(np.float32),
            "gates/bias": (1.0 + rng.uniform(-a, a, 2 * H)).astype(np.float32),
            "candidate/kernel": rng.uniform(-a, a, (D + H, H)).astype(np.float32),
            "candidate/bias": rng.uniform(-a, a, H).astype(np.float32)}


@pytest.mark.parametrize("T,B,D,H", [(7, 3, 5, 8), (12, 6, 20, 32), (1, 2, 4, 16), (9, 17, 24, 48)])
def test_bgru_layer_forward_backward(cuda, T, B, D, H):
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(T * 10 + B)
    P = {d: _params(rng, D, H) for d in ("fw", "bw")}
    x = rng.randn(T, B, D).astype(np.float32)
    seq = np.array([T] + [int(rng.randint(max(T // 2, 1), T + 1)) for _ in range(B - 1)], np.int32)
    dy = rng.randn(T, B, 2 * H).astype(np.float32)
    # oracle (fp64 autograd)
    vs = {d: {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P[d].items()} for d in P}
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    y_ref, (hf, hb) = olstm.gru_forward(xt.transpose(0, 1), seq, [vs], True)
    (y_ref * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    # CUDA
    dev = cuda
    Pd = {d: {k: torch.tensor(v, device=dev) for k, v in P[d].items()} for d in P}
    Gd = {d: {k: torch.zeros_like(v) for k, v in Pd[d].items()} for d in P}
    desc = ops.gru_desc(T, B, D, H)
    xd, sd = torch.tensor(x, device=dev), torch.tensor(seq, device=dev)
    y, fs, res = ops.bgru_layer_forward(desc, xd, sd, Pd["fw"], Pd["bw"], want_final_state=True)
    dx = ops.bgru_layer_backward(desc, xd, sd, Pd["fw"], Pd["bw"], torch.tensor(dy, device=dev), res, Gd["fw"], Gd["bw"])
    torch.cuda.synchronize()
    np.testing.assert_allclose(y.cpu().numpy(), y_ref.detach().numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(fs[0].cpu().numpy(), hf.detach().numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(fs[1].cpu().numpy(), hb.detach().numpy(), rtol=2e-4, atol=2e-5)
    g = xt.grad.numpy()
    np.testing.assert_allclose(dx.cpu().numpy(), g, rtol=0, atol=5e-4 * max(1e-3, np.abs(g).max()))
    for d in P:
        for k in P[d]:
            g = vs[d][k].grad.numpy()
            np.testing.assert_allclose(Gd[d][k].cpu().numpy(), g, rtol=0, atol=5e-4 * max(1e-3, np.abs(g).max()),
                                       err_msg="%s/%s" % (d, k))


@pytest.mark.parametrize("encoder_type", ["bgru", "gru"])
def test_gru_ctc_model(cuda, encoder_type):
    from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC
    rng = np.random.RandomState(21)
    B, T, D, H, L, C = 4, 16, 18, 24, 2, 8
    model = CTC(encoder_type=encoder_type, input_size=D, num_units=H, num_layers=L, num_classes=C, parameter_init=0.2,
                clip_grad_norm=5.0, device=cuda, seed=9)
    names = [v.name for v in model.trainable_variables()]
    assert ("bgru_hidden1/fw/gru_cell/gates/kernel" if encoder_type == "bgru" else
            "multi_gru/multi_rnn_cell/cell_0/gru_cell/gates/kernel") in names
    x = rng.randn(B, T, D).astype(np.float32)
    seq = np.array([T, 11, 16, 6], np.int32)
    for b in range(B):
        x[b, seq[b]:] = 0
    labels = [list(rng.randint(0, C, size=int(rng.randint(2, 5)))) for _ in range(B)]
    loss, logits = model.compute_loss(x, labels, seq, keep_prob=1.0)
    model._backward()
    torch.cuda.synchronize()
    vs = {v.name: torch.tensor(v.tensor.cpu().numpy(), dtype=torch.float64, requires_grad=True)
          for v in model.trainable_variables()}
    l_ref, logits_ref, _ = omodel.ctc_model_forward(vs, torch.tensor(x, dtype=torch.float64), seq, labels, L,
                                                    gru=encoder_type)
    l_ref.backward()
    assert abs(float(loss) - float(l_ref.detach())) <= 2e-4 * abs(float(l_ref.detach()))
    np.testing.assert_allclose(logits.cpu().numpy(), logits_ref.detach().numpy(), rtol=2e-4, atol=2e-5)
    for v in model.trainable_variables():
        g = vs[v.name].grad.numpy()
        np.testing.assert_allclose(v.grad.cpu().numpy(), g, rtol=0, atol=1e-3 * max(1e-4, np.abs(g).max()),
                                   err_msg=v.name)
    loss, _ = model.compute_loss(x, labels, seq, keep_prob=0.9)
    model.train(loss, "adam", 1e-2)
    loss2, _ = model.compute_loss(x, labels, seq, keep_prob=1.0, is_training=False)
    assert np.isfinite(float(loss2))
