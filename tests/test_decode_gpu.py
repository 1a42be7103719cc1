"""Greedy decoder / posteriors on the GPU vs the oracle: label indices bit-exact."""
import numpy as np
import pytest
import torch

from oracle import decode as odec

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,B,C", [(30, 4, 6), (300, 8, 62), (1000, 64, 29), (257, 3, 3001), (513, 2, 5)])
def test_greedy_bit_exact(cuda, T, B, C):
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(T + C)
    logits = rng.randn(T, B, C).astype(np.float32)
    # make the blank and repeats frequent so collapse/removal both matter
    logits[:, :, C - 1] += 1.5
    logits[1::2] = logits[0::2][: logits[1::2].shape[0]] + 0.01 * rng.randn(*logits[1::2].shape).astype(np.float32)
    seq = np.array([T] + [int(rng.randint(1, T + 1)) for _ in range(B - 1)], np.int32)
    lab, n = ops.ctc_greedy_decode(torch.tensor(logits, device=cuda), torch.tensor(seq, device=cuda))
    torch.cuda.synchronize()
    lab, n = lab.cpu().numpy(), n.cpu().numpy()
    ref = odec.greedy_decode(np.transpose(logits, (1, 0, 2)), seq, C - 1)
    for b in range(B):
        assert n[b] == len(ref[b])
        assert list(lab[b, :n[b]]) == ref[b]
        assert np.all(lab[b, n[b]:] == -1)


def test_greedy_ties_first_index(cuda):
    from tensorflow_end2end_speech_recognition_b200 import ops
    T, B, C = 6, 1, 40
    logits = np.zeros((T, B, C), np.float32)        # all ties -> argmax 0 everywhere
    logits[2, 0, 37] = 1.0
    logits[3, 0, 5] = 1.0
    logits[3, 0, 33] = 1.0                          # tie between 5 and 33 -> 5
    lab, n = ops.ctc_greedy_decode(torch.tensor(logits, device=cuda),
                                   torch.tensor([T], dtype=torch.int32, device=cuda))
    assert list(lab[0, :n[0]].cpu().numpy()) == [0, 37, 5, 0]


def test_softmax_rows(cuda):
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(0)
    x = (rng.randn(777, 29) * 3).astype(np.float32)
    y = ops.softmax_rows(torch.tensor(x, device=cuda)).cpu().numpy()
    e = np.exp(x.astype(np.float64) - x.max(-1, keepdims=True))
    np.testing.assert_allclose(y, e / e.sum(-1, keepdims=True), rtol=1e-5, atol=1e-7)


def _beam_gpu(probs, seq, beam, cuda):
    from tensorflow_end2end_speech_recognition_b200 import ops
    with np.errstate(divide="ignore"):
        lp = np.log(probs)                      # float32 log like the reference (np.log(probs))
    lab, n, sc = ops.ctc_beam_decode(torch.tensor(lp, device=cuda),
                                     torch.tensor(np.asarray(seq, np.int32), device=cuda), beam)
    torch.cuda.synchronize()
    lab, n, sc = lab.cpu().numpy(), n.cpu().numpy(), sc.cpu().numpy()
    return [list(lab[b, :n[b]]) for b in range(len(seq))], sc


def test_beam_search_matches_reference_golden_vectors(cuda):
    """bit-exact label indices against the vectors produced by the reference's own numpy
    BeamSearchDecoder (tests/golden/make_golden.py)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ctc_decoders.npz"))
    for i in range(int(z["n"])):
        p = z["probs_%d" % i]
        T = p.shape[1]
        labs, sc = _beam_gpu(p, [T], int(z["beam_%d" % i]), cuda)
        assert labs[0] == list(z["beam_labels_%d" % i]), "case %d" % i
        assert abs(sc[0] - float(z["beam_score_%d" % i])) < 1e-4 * max(1.0, abs(sc[0]))


@pytest.mark.parametrize("T,B,C,beam,peaky", [(40, 6, 29, 20, 3.0), (60, 4, 12, 5, 1.0), (100, 3, 29, 20, 6.0),
                                              (30, 5, 62, 10, 2.0), (25, 4, 5, 1, 1.0)])
def test_beam_search_random_vs_oracle(cuda, T, B, C, beam, peaky):
    rng = np.random.RandomState(T * 3 + C)
    x = rng.randn(B, T, C) * peaky
    x[..., C - 1] += 0.5 * peaky
    p = np.exp(x - x.max(-1, keepdims=True))
    p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
    seq = [T] + [int(rng.randint(T // 2, T + 1)) for _ in range(B - 1)]
    labs, sc = _beam_gpu(p, seq, beam, cuda)
    ref, rsc = odec.beam_search_decode(p, seq, C - 1, beam)
    assert labs == ref
    np.testing.assert_allclose(sc, np.asarray(rsc), rtol=1e-5, atol=1e-4)


def test_edit_distance_kernel_bit_exact(cuda):
    """b2_edit_distance vs the textbook DP on random pairs, incl. empty / long / identical rows."""
    from tensorflow_end2end_speech_recognition_b200 import ops

    def lev(a, b):
        d = list(range(len(b) + 1))
        for i in range(1, len(a) + 1):
            prev, d[0] = d[0], i
            for j in range(1, len(b) + 1):
                cur = d[j]
                d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (a[i - 1] != b[j - 1]))
                prev = cur
        return d[len(b)]
    rng = np.random.RandomState(0)
    hyp, ref = [], []
    for n, m in [(0, 5), (5, 0), (1, 1), (7, 7), (33, 31), (64, 65), (100, 250), (250, 100), (3, 40), (32, 32)]:
        hyp.append(list(rng.randint(0, 4, n)))
        ref.append(list(rng.randint(0, 4, m)))
    hyp.append(list(range(50))); ref.append(list(range(50)))
    for _ in range(20):
        hyp.append(list(rng.randint(0, 28, int(rng.randint(0, 120)))))
        ref.append(list(rng.randint(0, 28, int(rng.randint(1, 120)))))
    got = ops.edit_distance(hyp, ref, cuda)
    want = np.array([lev(a, b) for a, b in zip(hyp, ref)])
    assert np.array_equal(got, want)


def test_per_cer_wer_wrappers(cuda):
    """utils/evaluation/edit_distance.py:35-109 call shapes over the device Levenshtein kernel"""
    from tensorflow_end2end_speech_recognition_b200.utils.evaluation import edit_distance as ed

    def lev(a, b):
        d = list(range(len(b) + 1))
        for i in range(1, len(a) + 1):
            prev, d[0] = d[0], i
            for j in range(1, len(b) + 1):
                cur = d[j]
                d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (a[i - 1] != b[j - 1]))
                prev = cur
        return d[len(b)]
    ref = "she had your dark suit in greasy wash water all year".split()
    hyp = "she had dark suite in greasy wash water water all year".split()
    assert ed.compute_wer(ref, hyp, normalize=False) == lev(ref, hyp)
    assert abs(ed.compute_wer(ref, hyp) - lev(ref, hyp) / len(ref)) < 1e-12
    s, i, d = ed.wer_align(ref, hyp)
    assert s + i + d == lev(ref, hyp)
    a, b = "shehadyourdarksuit", "shehadyrdarksuite"
    assert abs(ed.compute_cer(b, a) - lev(list(b), list(a)) / len(a)) < 1e-12
    pr, ph = ["sh", "iy", "hv", "ae", "d"], ["sh", "ih", "hv", "d"]
    assert abs(ed.compute_per(pr, ph) - lev(pr, ph) / len(pr)) < 1e-12
    # batched form + the sparse-triple form (normalised by the prediction length, as the reference's swapped call does)
    from tensorflow_end2end_speech_recognition_b200.utils.io.labels.sparsetensor import list2sparsetensor
    import numpy as np
    true = np.array([[1, 2, 3, 4], [5, 6, -1, -1]]); pred = np.array([[1, 3, 4, -1], [5, 6, 7, -1]])
    out = ed.compute_edit_distance(None, list2sparsetensor(true, -1), list2sparsetensor(pred, -1))
    np.testing.assert_allclose(out, [1 / 3, 1 / 3], rtol=1e-6)


@pytest.mark.parametrize("T,B,C,W", [(30, 4, 6, 3), (60, 5, 29, 20), (40, 3, 29, 100), (25, 2, 301, 16),
                                     (12, 2, 3001, 100), (1, 2, 5, 4)])
@pytest.mark.parametrize("merge", [True, False])
def test_tf_semantics_beam_search(cuda, T, B, C, W, merge):
    """b2_ctc_beam_decode_tf vs the restatement of tf.nn.ctc_beam_search_decoder (oracle/decode.py): label sequences
    identical, scores equal to fp32 rounding; peaky and flat posteriors, ragged lengths, merge_repeated on/off"""
    import torch
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(T * 131 + C + W)
    logits = (rng.randn(T, B, C) * (3.0 if C < 100 else 1.5)).astype(np.float32)
    logits[:, :, C - 1] += 1.0                       # blanks a little more likely, as in a trained model
    seq = np.array([T] + [int(rng.randint(max(1, T // 2), T + 1)) for _ in range(B - 1)], np.int32)
    lab, n, score = ops.ctc_beam_decode_tf(torch.tensor(logits, device=cuda), torch.tensor(seq, device=cuda), W,
                                           blank=C - 1, merge_repeated=merge)
    lab, n, score = lab.cpu().numpy(), n.cpu().numpy(), score.cpu().numpy()
    for b in range(B):
        paths, scores = odec.tf_ctc_beam_search_single(logits[:seq[b], b], C - 1, W, merge_repeated=merge)
        assert list(lab[b, :n[b]]) == paths[0], (b, list(lab[b, :n[b]]), paths[0])
        assert (lab[b, n[b]:] == -1).all()
        assert abs(score[b] - scores[0]) <= 1e-4 * max(1.0, abs(scores[0]))


def test_tf_beam_width_one_is_not_greedy_but_close(cuda):
    """model-level: CTC.decoder(beam_width>1) takes the TF-semantics path and merges repeats in the output"""
    import torch
    from tensorflow_end2end_speech_recognition_b200 import ops
    T, C = 6, 4
    logits = np.full((T, 1, C), -5.0, np.float32)
    for t, c in enumerate([0, 3, 0, 1, 3, 1]):        # a - a b - b  -> "a a b b"
        logits[t, 0, c] = 5.0
    seq = np.array([T], np.int32)
    for merge, want in ((True, [0, 1]), (False, [0, 0, 1, 1])):
        lab, n, _ = ops.ctc_beam_decode_tf(torch.tensor(logits, device=cuda), torch.tensor(seq, device=cuda), 8,
                                           blank=C - 1, merge_repeated=merge)
        assert list(lab.cpu().numpy()[0, :int(n[0])]) == want
