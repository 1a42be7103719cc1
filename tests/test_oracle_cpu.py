"""Pins the CPU oracle (test infrastructure) against everything this environment
offers: the reference's own numpy decoders (golden vectors), brute-force path
enumeration, torch's independent CTC, and a second literal LSTM form."""
import os

import numpy as np
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
