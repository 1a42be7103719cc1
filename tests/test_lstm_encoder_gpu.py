"""encoder_type 'lstm' / 'vgg_lstm' (reference: models/encoders/core/lstm.py, vgg_lstm.py): the unidirectional stack
runs on the BLSTM layer kernels with an idle (all-zero) second direction; CTC loss, logits and every gradient vs the
fp64 oracle's unidirectional stack (oracle/lstm.py::lstm_forward)."""
import numpy as np
import pytest
import torch

from oracle import model as omodel

pytestmark = pytest.mark.gpu


def _check(model, x, seq, labels, L, tol_loss, tol_logits, tol_grad, vgg=None):
    loss, logits = model.compute_loss(x, labels, seq, keep_prob=1.0)
    model._backward()
    torch.cuda.synchronize()
    vs = {v.name: torch.tensor(v.tensor.cpu().numpy(), dtype=torch.float64, requires_grad=True)
          for v in model.trainable_variables()}
    l_ref, logits_ref, _ = omodel.ctc_model_forward(vs, torch.tensor(x, dtype=torch.float64), seq, labels, L,
                                                    vgg=vgg, unidirectional=True)
    l_ref.backward()
    assert abs(float(loss) - float(l_ref.detach())) <= tol_loss * abs(float(l_ref.detach()))
    lg, lr = logits.cpu().numpy(), logits_ref.detach().numpy()
    assert lg.shape == lr.shape
    assert np.abs(lg - lr).max() <= tol_logits * max(1.0, np.abs(lr).max())
    for v in model.trainable_variables():
        g = vs[v.name].grad.numpy()
        s = max(1e-4, np.abs(g).max())
        np.testing.assert_allclose(v.grad.cpu().numpy(), g, rtol=0, atol=tol_grad * s, err_msg=v.name)


@pytest.mark.parametrize("precision,tols", [("fp32", (2e-4, 2e-4, 1e-3)), ("bf16", (1e-2, 3e-2, 6e-2))])
def test_lstm_ctc_model(cuda, precision, tols):
    from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC
    rng = np.random.RandomState(11)
    B, T, D, H, L, C = 5, 18, 24, 64, 3, 10
    model = CTC(encoder_type="lstm", input_size=D, num_units=H, num_layers=L, num_classes=C, parameter_init=0.1,
                clip_grad_norm=5.0, precision=precision, device=cuda, seed=5)
    names = [v.name for v in model.trainable_variables()]
    assert "multi_lstm/multi_rnn_cell/cell_0/lstm_cell/kernel" in names and model.encoder.output_size == H
    assert model.variables["multi_lstm/multi_rnn_cell/cell_1/lstm_cell/kernel"].shape == (2 * H, 4 * H)
    x = rng.randn(B, T, D).astype(np.float32)
    seq = np.array([T, 12, 18, 7, 15], np.int32)
    for b in range(B):
        x[b, seq[b]:] = 0
    labels = [list(rng.randint(0, C, size=int(rng.randint(2, 6)))) for _ in range(B)]
    _check(model, x, seq, labels, L, *tols)
    loss, _ = model.compute_loss(x, labels, seq, keep_prob=1.0)
    model.train(loss, "adam", 1e-2)
    loss2, _ = model.compute_loss(x, labels, seq, keep_prob=1.0, is_training=False)
    assert float(loss2) < float(loss)


def test_vgg_lstm_ctc_model(cuda):
    from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC
    rng = np.random.RandomState(12)
    B, T, nch, C, H, L = 3, 14, 8, 9, 32, 2
    D = nch * 3
    model = CTC(encoder_type="vgg_lstm", input_size=D, num_units=H, num_layers=L, num_classes=C,
                parameter_init=0.1, clip_grad_norm=5.0, device=cuda, seed=6)
    x = rng.randn(B, T, D).astype(np.float32)
    seq = np.array([T, 9, 12], np.int32)
    for b in range(B):
        x[b, seq[b]:] = 0
    labels = [list(rng.randint(0, C, size=int(rng.randint(2, 5)))) for _ in range(B)]
    _check(model, x, seq, labels, L, 2e-4, 2e-4, 1e-3, vgg=(nch, 1))


def test_load_encoder_registry():
    from tensorflow_end2end_speech_recognition_b200.models.encoders.load_encoder import load
    assert load("lstm").__name__ == "LSTMEncoder" and load("vgg_lstm").__name__ == "VGGLSTMEncoder"
    with pytest.raises(ValueError):
        load("cnn_zhang")
