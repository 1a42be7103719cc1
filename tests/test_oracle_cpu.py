"""Pins the CPU oracle (test infrastructure) against everything this environment
offers: the reference's own numpy decoders (golden vectors), brute-force path
enumeration, torch's independent CTC, and a second literal LSTM form."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc as octc
from oracle import decode as odec
from oracle import lstm as olstm
from oracle import optim as oopt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ctc_decoders.npz")


def test_decoders_match_reference_golden_vectors():
    z = np.load(GOLD)
    for i in range(int(z["n"])):
        p = z["probs_%d" % i]
        T, C = p.shape[1], p.shape[2]
        assert odec.greedy_decode(p, [T], C - 1)[0] == list(z["greedy_%d" % i])
        lab, sc = odec.beam_search_decode(p, [T], C - 1, int(z["beam_%d" % i]))
        assert lab[0] == list(z["beam_labels_%d" % i])
        # scores: the reference mixes float32/float64 scalars (numpy-version dependent)
        assert abs(sc[0] - float(z["beam_score_%d" % i])) < 1e-6 * abs(sc[0])


def test_ctc_brute_force():
    rng = np.random.RandomState(0)
    for labels in ([1], [0, 1], [1, 1], [2, 0, 2], []):
        x = rng.randn(6, 4)
        nll = octc.ctc_loss_single(x, labels, 3)[0]
        assert abs(nll - octc.ctc_brute_force(x, labels, 3)) < 1e-10


def test_ctc_matches_torch_and_fast_form():
    rng = np.random.RandomState(1)
    T, B, C = 25, 4, 7
    logits = rng.randn(T, B, C)
    labels = [[1, 1, 2], [0], [3, 4, 5, 5, 0], [2, 2]]
    seq = [25, 11, 25, 9]
    l1, g1 = octc.ctc_loss(logits, labels, seq)
    l2, g2 = octc.ctc_loss_fast(logits, labels, seq)
    np.testing.assert_allclose(l1, l2, rtol=1e-12)
    np.testing.assert_allclose(g1, g2, atol=1e-12)
    x = torch.tensor(logits, requires_grad=True)
    tl = torch.nn.functional.ctc_loss(torch.log_softmax(x, -1), torch.tensor(sum(labels, [])),
                                      torch.tensor(seq), torch.tensor([len(l) for l in labels]),
                                      blank=C - 1, reduction="none")
    tl.sum().backward()
    np.testing.assert_allclose(l1, tl.detach().numpy(), rtol=1e-10)
    np.testing.assert_allclose(g1, x.grad.numpy(), atol=1e-10)


def test_ctc_skip_and_error_semantics():
    x = np.random.RandomState(2).randn(4, 2, 5)
    loss, grad = octc.ctc_loss(x, [[1, 2, 3, 1, 2], [1]], [4, 4], ignore_longer_outputs_than_inputs=True)
    assert loss[0] == 0 and np.all(grad[:, 0] == 0) and loss[1] > 0
    try:
        octc.ctc_loss(x, [[1, 2, 3, 1, 2], [1]], [4, 4], ignore_longer_outputs_than_inputs=False)
        assert False
    except ValueError:
        pass


def test_lstm_two_forms_and_masking():
    L = olstm.init_blstm_params(6, 8, 2, seed=1, dtype=np.float64)
    rng = np.random.RandomState(0)
    x = rng.randn(3, 7, 6)
    seq = [7, 4, 5]
    y, fs = olstm.blstm_forward(torch.tensor(x), seq, L)
    y2 = olstm.blstm_forward_numpy(x, seq, L)
    np.testing.assert_allclose(y.numpy(), y2, atol=1e-12)
    # outputs are zero past the length; fw final h equals the output at len-1, bw at t=0
    assert np.all(y.numpy()[4:, 1] == 0)
    (cf, hf), (cb, hb) = fs
    np.testing.assert_allclose(hf[1].numpy(), y.numpy()[3, 1, :8], atol=1e-12)
    np.testing.assert_allclose(hb[1].numpy(), y.numpy()[0, 1, 8:], atol=1e-12)
    # padding content must not matter
    x2 = x.copy()
    x2[1, 4:] = 99.0
    y3, _ = olstm.blstm_forward(torch.tensor(x2), seq, L)
    np.testing.assert_allclose(y.numpy(), y3.numpy(), atol=1e-12)


def test_optimizers_against_torch():
    rng = np.random.RandomState(3)
    w0 = rng.randn(20)
    grads = [rng.randn(20) for _ in range(5)]
    pairs = {"sgd": torch.optim.SGD, "momentum": torch.optim.SGD, "nestrov": torch.optim.SGD,
             "adam": torch.optim.Adam}
    for name, cls in pairs.items():
        w = torch.tensor(w0.copy(), requires_grad=True)
        kw = {"lr": 0.01}
        if name == "momentum":
            kw["momentum"] = 0.9
        if name == "nestrov":
            kw.update(momentum=0.9, nesterov=True)
        topt = cls([w], **kw)
        ref = [w0.copy()]
        o = oopt.Optimizer(name, 0.01)
        for g in grads:
            w.grad = torch.tensor(g.copy())
            topt.step()
            o.step(ref, [g])
        tol = 1e-6 if name == "adam" else 1e-12     # Adam: eps placement differs (TF: epsilon-hat)
        np.testing.assert_allclose(ref[0], w.detach().numpy(), rtol=tol, atol=tol)


def test_clip_and_average():
    g = np.array([3.0, 4.0])
    np.testing.assert_allclose(oopt.clip_by_norm(g, 1.0), [0.6, 0.8])
    np.testing.assert_allclose(oopt.clip_by_norm(g, 10.0), g)
    avg = oopt.average_gradients([[np.ones(3), None], [3 * np.ones(3), np.ones(2)]])
    np.testing.assert_allclose(avg[0], 2 * np.ones(3))
    np.testing.assert_allclose(avg[1], np.ones(2))


def test_attention_torch_twin_matches_numpy_oracle():
    """oracle/seq2seq.py (torch, autograd) restates oracle/attention.py and
    oracle/attention_decoder.py; the two forms must agree."""
    import torch
    from oracle import attention as oatt, attention_decoder as odec, seq2seq as os2s
    rng = np.random.RandomState(0)
    B, T, E, Dq, A = 3, 9, 8, 6, 5
    for t in oatt.ATTENTION_TYPE:
        p = {"W_query/weights": rng.randn(Dq, A), "W_keys/weights": rng.randn(E, A if t != "luong_general" else Dq),
             "W_keys/biases": rng.randn(A), "filter": rng.randn(5, 1, 10), "W_filter/weights": rng.randn(10, A),
             "W_filter/biases": rng.randn(A), "v_a": rng.randn(A), "W_concat/weights": rng.randn(E + Dq, A)}
        enc = rng.randn(B, T, E)
        q = rng.randn(B, E if t == "luong_dot" else Dq)
        lens = np.array([9, 4, 6])
        pa = np.abs(rng.rand(B, T))
        for sig in (False, True):
            a, c = oatt.attention_step(t, enc, q, lens, pa, p, 1.5, sig)
            pt = {k: torch.tensor(v) for k, v in p.items()}
            a2, c2 = os2s.attention_step_t(t, torch.tensor(enc), torch.tensor(q), lens, torch.tensor(pa), pt, 1.5, sig)
            np.testing.assert_allclose(a, a2.numpy(), atol=1e-12)
            np.testing.assert_allclose(c, c2.numpy(), atol=1e-12)
    # teacher-forced decoder loop, both forms
    Hd, emb, C = 6, 4, 7
    p = {"W_embedding": rng.randn(C, emb), "attentional_vector/weights": rng.randn(Hd + E, Hd) * .3,
         "output_layer/weights": rng.randn(Hd, C), "output_layer/biases": rng.randn(C) * .1,
         "cell": {"kernel": rng.randn(emb + E + Hd, 4 * Hd) * .3, "bias": rng.randn(4 * Hd) * .1,
                  "w_i_diag": rng.randn(Hd) * .1, "w_f_diag": rng.randn(Hd) * .1, "w_o_diag": rng.randn(Hd) * .1},
         "attention": {"W_keys/weights": rng.randn(E, Hd) * .3}}
    enc = rng.randn(B, T, E)
    lens = np.array([9, 3, 5])
    labels = np.array([[5, 1, 2, 3, 6], [5, 0, 6, 6, 6], [5, 1, 1, 6, 6]])
    lab_len = np.array([5, 3, 4])
    st = (rng.randn(B, Hd) * .1, rng.randn(B, Hd) * .1)
    r = odec.decode(p, "luong_general", enc, lens, st, labels=labels, labels_seq_len=lab_len)
    tt = lambda x: {k: tt(v) for k, v in x.items()} if isinstance(x, dict) else torch.tensor(x)
    r2 = os2s.decode_train_t(tt(p), "luong_general", torch.tensor(enc), lens, (torch.tensor(st[0]), torch.tensor(st[1])),
                             labels, lab_len)
    np.testing.assert_allclose(r["logits"], r2["logits"].numpy(), atol=1e-12)
    assert np.array_equal(r["predicted_ids"], r2["predicted_ids"].numpy())
    # sequence loss: hand-computed masked mean
    logits = rng.randn(2, 3, 4)
    tg = np.array([[1, 2, 0], [3, 0, 0]])
    ln = np.array([3, 1])
    lp = logits - np.log(np.exp(logits).sum(-1, keepdims=True))
    want = -(lp[0, 0, 1] + lp[0, 1, 2] + lp[0, 2, 0] + lp[1, 0, 3]) / 4.0
    assert abs(float(os2s.sequence_loss_t(torch.tensor(logits), tg, ln)) - want) < 1e-12


def test_greedy_decoder_oracle_semantics():
    """dynamic_decode rules: stop when every row emitted <EOS> (or at max length), finished rows
    emit zeros and keep their state."""
    from oracle import attention_decoder as odec
    rng = np.random.RandomState(1)
    B, T, E, Hd, emb, C = 3, 7, 8, 6, 4, 5
    p = {"W_embedding": rng.randn(C, emb), "attentional_vector/weights": rng.randn(Hd + E, Hd) * .3,
         "output_layer/weights": rng.randn(Hd, C), "output_layer/biases": np.zeros(C),
         "cell": {"kernel": rng.randn(emb + E + Hd, 4 * Hd) * .3, "bias": np.zeros(4 * Hd)},
         "attention": {"W_keys/weights": rng.randn(E, Hd) * .3}}
    enc = rng.randn(B, T, E)
    lens = np.array([7, 3, 5])
    st = (np.zeros((B, Hd)), np.zeros((B, Hd)))
    r = odec.decode(p, "luong_general", enc, lens, st, sos=3, eos=4, max_decode_length=6)
    ids = r["predicted_ids"]
    assert ids.shape[1] <= 6
    for b in range(B):
        hit = np.where(ids[b] == 4)[0]
        if len(hit):
            assert np.all(r["logits"][b, hit[0] + 1:] == 0) and np.all(ids[b, hit[0] + 1:] == 0)
    p["output_layer/biases"] = np.array([0, 0, 0, 0, 50.0])
    r = odec.decode(p, "luong_general", enc, lens, st, sos=3, eos=4, max_decode_length=6)
    assert r["logits"].shape[1] == 1


def test_input_pipeline_oracle_against_reference_golden():
    """oracle/inputs.py vs the outputs of the reference's own stack_frame / do_splice."""
    import os
    from oracle import inputs as oin
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "input_pipeline.npz"))
    for i in range(int(g["n_cases"])):
        T, nch, S, K, P = [int(v) for v in g["cfg_%d" % i]]
        st = oin.stack_frame(g["x_%d" % i], S, K)
        assert np.array_equal(st, g["stacked_%d" % i])
        assert np.array_equal(oin.do_splice(st, P, S), g["spliced_%d" % i])
    with pytest.raises(ValueError):
        oin.stack_frame(np.zeros((4, 3), np.float32), 2, 3)
    ins, labs, lens = oin.make_batch([np.ones((5, 3), np.float32), np.ones((2, 3), np.float32), np.ones((4, 3), np.float32)],
                                     [[1, 2], [3], [4, 5, 6]], num_gpu=2)
    assert [a.shape for a in ins] == [(2, 5, 3), (1, 5, 3)] and labs[0].tolist() == [[1, 2, -1], [3, -1, -1]]
    assert lens[1].tolist() == [4]


def test_beam_search_oracle_width_one_is_greedy():
    """oracle beam search with one beam and no length penalty follows the arg-max path of the greedy oracle."""
    from oracle import attention_decoder as odec
    rng = np.random.RandomState(2)
    T, E, Hd, emb, C = 9, 8, 6, 4, 7
    p = {"W_embedding": rng.randn(C, emb), "attentional_vector/weights": rng.randn(Hd + E, Hd) * .3,
         "output_layer/weights": rng.randn(Hd, C), "output_layer/biases": rng.randn(C) * .1,
         "cell": {"kernel": rng.randn(emb + E + Hd, 4 * Hd) * .3, "bias": np.zeros(4 * Hd)},
         "attention": {"W_keys/weights": rng.randn(E, Hd) * .3}}
    enc = rng.randn(1, T, E)
    st = (rng.randn(1, Hd) * .1, rng.randn(1, Hd) * .1)
    g = odec.decode(p, "luong_general", enc, np.array([T]), st, sos=5, eos=6, max_decode_length=8)
    b = odec.beam_search_decode(p, "luong_general", enc[0], T, (st[0][0], st[1][0]), 5, 6, 1, 0.0, 8)
    gi = g["predicted_ids"][0]
    n = min(len(gi), b["ids"].shape[1])
    assert np.array_equal(b["ids"][0, :n], gi[:n])
    # wider beams never score worse than the greedy path
    b4 = odec.beam_search_decode(p, "luong_general", enc[0], T, (st[0][0], st[1][0]), 5, 6, 4, 0.0, 8)
    assert b4["log_probs"].max() >= b["log_probs"][0] - 1e-9


def test_vgg_oracle_two_forms_and_geometry():
    """oracle/vgg.py: the torch conv / pool form against literal numpy loops (SAME padding, pooling tails)."""
    import torch
    from oracle import vgg as ovgg
    rng = np.random.RandomState(0)
    for (H, W) in ((7, 1), (6, 3), (5, 4), (1, 1)):
        x = rng.randn(2, H, W, 3)
        w = rng.randn(3, 3, 3, 5)
        b = rng.randn(5)
        a = ovgg._conv_relu(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
        np.testing.assert_allclose(a, ovgg.conv_relu_numpy(x, w, b), atol=1e-12)
        np.testing.assert_allclose(ovgg._pool(torch.tensor(a)).numpy(), ovgg.max_pool_numpy(a), atol=0)
    assert ovgg.output_geometry(80, 1) == (20, 1) and ovgg.output_geometry(7, 3) == (2, 1)
    # whole front-end: shapes and a 1-wide image never touches the side columns of the filters
    nch, Wd = 6, 1
    p = {}
    chans = (3, 64, 64, 128, 128)
    for i, n in enumerate(ovgg.CONVS):
        p[n + "/weight"] = torch.tensor(rng.randn(3, 3, chans[i], chans[i + 1]) * 0.05, requires_grad=True)
        p[n + "/bias"] = torch.zeros(chans[i + 1], dtype=torch.float64)
    h4, w4 = ovgg.output_geometry(nch, Wd)
    p["bridge/weights"] = torch.tensor(rng.randn(h4 * w4 * 128, 256) * 0.05)
    p["bridge/biases"] = torch.zeros(256, dtype=torch.float64)
    out = ovgg.vgg_frontend(torch.tensor(rng.randn(2, 3, nch * Wd * 3)), p, nch, Wd)
    assert out.shape == (2, 3, 256)
    out.sum().backward()
    g = p["VGG1/conv2/weight"].grad.numpy()
    assert np.all(g[:, 0] == 0) and np.all(g[:, 2] == 0) and np.abs(g[:, 1]).max() > 0


def test_seq2seq_oracle_loss_gradient_matches_finite_differences():
    """oracle/seq2seq.py end to end on CPU (encoder -> bridge -> teacher-forced decoder -> sequence loss, and the
    joint CTC-attention total): autograd gradient of a few parameters vs central differences."""
    import torch
    from oracle import lstm as olstm, seq2seq as os2s
    rng = np.random.RandomState(3)
    B, T, D, H, Hd, A, emb, V = 2, 6, 4, 3, 5, 4, 3, 4
    C = V + 2
    layers = olstm.init_blstm_params(D, H, 1, parameter_init=0.3, seed=1, dtype=np.float64)
    vs = {}
    for d in ("fw", "bw"):
        for k, v in layers[0][d].items():
            vs["encoder/blstm_hidden1/%s/lstm_cell/%s" % (d, k)] = v
    r = lambda *s: rng.randn(*s) * 0.3
    vs.update({"decoder/bridge/weights": r(4 * H, 2 * Hd), "decoder/bridge/biases": r(2 * Hd),
               "decoder/output_embedding/W_embedding": r(C, emb),
               "decoder/decoder_rnn_cell/lstm_cell/kernel": r(emb + 2 * H + Hd, 4 * Hd),
               "decoder/decoder_rnn_cell/lstm_cell/bias": r(4 * Hd),
               "decoder/attention_decoder/attention_layer/W_query/weights": r(Hd, A),
               "decoder/attention_decoder/attention_layer/W_keys/weights": r(2 * H, A),
               "decoder/attention_decoder/attention_layer/W_keys/biases": r(A),
               "decoder/attention_decoder/attention_layer/v_a": r(A),
               "decoder/attention_decoder/attentional_vector/weights": r(Hd + 2 * H, Hd),
               "decoder/attention_decoder/output_layer/weights": r(Hd, C),
               "decoder/attention_decoder/output_layer/biases": r(C),
               "ctc_output/weights": r(2 * H, V + 1), "ctc_output/biases": r(V + 1)})
    x = rng.randn(B, T, D)
    seq = np.array([6, 4])
    labels = np.array([[V, 1, 2, V + 1], [V, 3, V + 1, V + 1]])
    lab_len = np.array([4, 3])
    ctc_labels = [[1, 2], [3]]
    cfg = dict(num_layers=1, attention_type="bahdanau_content", lambda_weight=0.3, weight_decay=1e-2)

    def loss_of(values):
        t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in values.items()}
        out = os2s.seq2seq_loss(t, cfg, torch.tensor(x), seq, labels, lab_len, ctc_labels)
        return out["total_loss"], t
    loss, t = loss_of(vs)
    loss.backward()
    for name, idx in (("decoder/attention_decoder/attention_layer/v_a", (1,)),
                      ("encoder/blstm_hidden1/bw/lstm_cell/kernel", (2, 5)),
                      ("decoder/bridge/weights", (3, 4)), ("ctc_output/weights", (1, 2)),
                      ("decoder/output_embedding/W_embedding", (1, 0))):
        eps = 1e-6
        vp = {k: v.copy() for k, v in vs.items()}
        vm = {k: v.copy() for k, v in vs.items()}
        vp[name][idx] += eps
        vm[name][idx] -= eps
        fd = (float(loss_of(vp)[0].detach()) - float(loss_of(vm)[0].detach())) / (2 * eps)
        assert abs(fd - float(t[name].grad[idx])) <= 1e-6 + 1e-5 * abs(fd), (name, fd, float(t[name].grad[idx]))


def test_unidirectional_oracle_is_the_forward_half_of_the_blstm_oracle():
    """oracle/lstm.py::lstm_forward (encoder_type 'lstm') against blstm_forward's forward direction, one layer, ragged."""
    import torch
    from oracle import lstm as ol
    layers = ol.init_blstm_params(6, 8, 1, parameter_init=0.3, seed=1)
    x = torch.tensor(np.random.RandomState(0).randn(3, 7, 6))
    seq = [7, 4, 6]
    both = [{d: {k: torch.tensor(v, dtype=torch.float64) for k, v in layers[0][d].items()} for d in layers[0]}]
    y2, fs2 = ol.blstm_forward(x, seq, both)
    y1, st = ol.lstm_forward(x, seq, [both[0]["fw"]])
    assert y1.shape == (7, 3, 8) and len(st) == 1
    assert float((y1 - y2[:, :, :8]).abs().max()) == 0.0
    assert float((st[0][0] - fs2[0][0]).abs().max()) == 0.0 and float((st[0][1] - fs2[0][1]).abs().max()) == 0.0
