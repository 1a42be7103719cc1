"""Host-side checks of the encoder registry and the TF variable names / shapes the encoder mirrors create (no GPU):
models/encoders/load_encoder.py:26-57, core/{blstm,lstm,gru,vgg_blstm,vgg_lstm,multitask_*}.py."""
import numpy as np
import pytest


def test_registry_matches_the_built_encoders():
    from tensorflow_end2end_speech_recognition_b200.models.encoders.load_encoder import ENCODERS, load
    assert sorted(ENCODERS) == ["bgru", "blstm", "gru", "lstm", "multitask_blstm", "multitask_lstm", "vgg_blstm",
                                "vgg_lstm"]
    for k, cls in ENCODERS.items():
        assert load(k) is cls
    with pytest.raises(ValueError):
        load("pyramid_blstm")


def _names(enc, d_in):
    return {n: a.shape for n, a in enc.create_variables(d_in, np.random.RandomState(0))}


def test_variable_names_and_shapes():
    from tensorflow_end2end_speech_recognition_b200.models.encoders.load_encoder import load
    D, H = 24, 16
    v = _names(load("blstm")(H, None, 2, "LSTMBlockCell", True, 0.1, None), D)
    assert v["blstm_hidden1/fw/lstm_cell/kernel"] == (D + H, 4 * H) and v["blstm_hidden2/bw/lstm_cell/kernel"] == (3 * H, 4 * H)
    assert v["blstm_hidden1/bw/lstm_cell/w_f_diag"] == (H,)
    v = _names(load("lstm")(H, None, 3, "LSTMBlockCell", True, 0.1, None), D)
    assert v["multi_lstm/multi_rnn_cell/cell_0/lstm_cell/kernel"] == (D + H, 4 * H)
    assert v["multi_lstm/multi_rnn_cell/cell_2/lstm_cell/kernel"] == (2 * H, 4 * H) and len(v) == 3 * 5
    v = _names(load("lstm")(H, None, 1, "BasicLSTMCell", True, 0.1, None), D)
    assert sorted(k.rsplit("/", 1)[1] for k in v) == ["bias", "kernel"]                 # no peepholes
    v = _names(load("bgru")(H, 2, 0.1), D)
    assert v["bgru_hidden1/fw/gru_cell/gates/kernel"] == (D + H, 2 * H)
    assert v["bgru_hidden2/bw/gru_cell/candidate/kernel"] == (3 * H, H) and len(v) == 2 * 2 * 4
    g = dict(load("bgru")(H, 1, 0.1).create_variables(D, np.random.RandomState(0)))
    assert np.all(g["bgru_hidden1/fw/gru_cell/gates/bias"] == 1.0) and np.all(g["bgru_hidden1/fw/gru_cell/candidate/bias"] == 0)
    v = _names(load("gru")(H, 2, 0.1), D)
    assert v["multi_gru/multi_rnn_cell/cell_1/gru_cell/gates/kernel"] == (2 * H, 2 * H) and len(v) == 2 * 4
    enc = load("vgg_lstm")(input_size=24, splice=1, num_stack=1, num_units=H, num_proj=None, num_layers=2,
                           lstm_impl="LSTMBlockCell", use_peephole=True, parameter_init=0.1, clip_activation=None)
    v = _names(enc, 24)
    assert "bridge/weights" in v and v["multi_lstm/multi_rnn_cell/cell_0/lstm_cell/kernel"] == (256 + H, 4 * H)
    assert enc.output_size == H
    ml = load("multitask_lstm")(H, None, 3, 2, "LSTMBlockCell", True, 0.1, None)
    assert ml.num_layers == 3 and ml.num_layers_sub == 2 and ml.output_size == H
    with pytest.raises(ValueError):
        load("multitask_lstm")(H, None, 2, 3, "LSTMBlockCell", True, 0.1, None)


def test_gru_oracle_unidirectional_is_the_forward_half():
    import torch
    from oracle import lstm as ol
    rng = np.random.RandomState(3)
    D, H, T, B = 5, 6, 8, 3
    p = {d: {"gates/kernel": torch.tensor(rng.randn(D + H, 2 * H) * 0.3), "gates/bias": torch.ones(2 * H, dtype=torch.float64),
             "candidate/kernel": torch.tensor(rng.randn(D + H, H) * 0.3), "candidate/bias": torch.zeros(H, dtype=torch.float64)}
         for d in ("fw", "bw")}
    x = torch.tensor(rng.randn(B, T, D))
    seq = [8, 5, 7]
    y2, (hf, hb) = ol.gru_forward(x, seq, [p], True)
    y1, h1 = ol.gru_forward(x, seq, [{"fw": p["fw"]}], False)
    assert float((y1 - y2[:, :, :H]).abs().max()) == 0.0 and float((h1 - hf).abs().max()) == 0.0
    assert float(y2[6:, 1].abs().max()) == 0.0                      # outputs are zero past the length
    # an all-zero idle direction stays exactly zero (the trick behind the unidirectional encoders)
    z = {k: torch.zeros_like(v) for k, v in p["bw"].items()}
    yz, (_, hz) = ol.gru_forward(x, seq, [{"fw": p["fw"], "bw": z}], True)
    assert float(yz[:, :, H:].abs().max()) == 0.0 and float(hz.abs().max()) == 0.0
