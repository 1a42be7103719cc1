"""CTC greedy / prefix-beam decoders, CPU restatement.  TEST INFRASTRUCTURE.

Greedy follows ``tf.nn.ctc_greedy_decoder`` as called at ``models/ctc/ctc.py:340-342``
(per-frame argmax over ``t < seq_len``, first index on ties, collapse repeats,
drop blanks) which is also what the reference's importable numpy decoder does
(``models/ctc/decoders/greedy_decoder.py:19-50``).

Beam search follows the reference's own numpy prefix beam search
``models/ctc/decoders/beam_search_decoder.py:53-152`` -- the decoder the
LibriSpeech evaluation script hard-selects (``examples/librispeech/evaluation/
eval_ctc.py:33-36,126-131``, beam_width=20 = BASELINE config 5).  Semantics that
matter for bit-exactness and are reproduced here:

* candidates = every (prefix in beam) x (class c), evaluated class-major,
  beam-minor (:92-134); blank keeps the prefix, a repeated last char updates
  both the extended prefix (only from p_b) and the unchanged prefix (from p_nb);
* ranking key = logsumexp(p_b, p_nb), sorted with Python's *stable* sort in
  descending order (:138-141), so ties keep dict insertion order;
* no pruning by probability, no output merge_repeated (unlike TF's
  ``ctc_beam_search_decoder(merge_repeated=True)``).

Arithmetic is float64 on ``log(probs)``, like the reference.
"""
import math
from collections import OrderedDict

import numpy as np

NEG_INF = -float("inf")


def greedy_decode(logits_or_probs_btc, seq_len, blank):
    """[B,T,C] scores (any monotone transform of probs) -> list of label lists."""
    out = []
    for b in range(len(seq_len)):
        tb = int(seq_len[b])
        idx = np.argmax(logits_or_probs_btc[b, :tb], axis=-1) if tb > 0 else np.zeros(0, np.int64)
        hyp = []
        prev = -1
        for k in idx:
            k = int(k)
            if k != prev and k != blank:
                hyp.append(k)
            prev = k
        out.append(hyp)
    return out


def _lse2(a, b):
    if a == NEG_INF and b == NEG_INF:
        return NEG_INF
    m = a if a > b else b
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def _lse3(a, b, c):
    if a == NEG_INF and b == NEG_INF and c == NEG_INF:
        return NEG_INF
    m = max(a, b, c)
    return m + math.log(math.exp(a - m) + math.exp(b - m) + math.exp(c - m))


def beam_search_decode_single(log_probs_tc, blank, beam_width):
    """One utterance, log-probabilities [T,C] (float64).  Returns (labels, neg_log_score)."""
    T, C = log_probs_tc.shape
    beam = [(tuple(), (0.0, NEG_INF))]
    for t in range(T):
        nxt = OrderedDict()

        def get(pfx):
            return nxt.get(pfx, (NEG_INF, NEG_INF))
        for c in range(C):
            p_t = float(log_probs_tc[t, c])
            for prefix, (p_b, p_nb) in beam:
                if c == blank:
                    nb, nnb = get(prefix)
                    nxt[prefix] = (_lse3(nb, p_b + p_t, p_nb + p_t), nnb)
                    continue
                end = prefix[-1] if prefix else None
                new_prefix = prefix + (c,)
                nb, nnb = get(new_prefix)
                if c != end:
                    nnb = _lse3(nnb, p_b + p_t, p_nb + p_t)
                else:
                    nnb = _lse2(nnb, p_b + p_t)
                nxt[new_prefix] = (nb, nnb)
                if c == end:
                    nb, nnb = get(prefix)
                    nxt[prefix] = (nb, _lse2(nnb, p_nb + p_t))
        beam = sorted(nxt.items(), key=lambda kv: _lse2(*kv[1]), reverse=True)[:beam_width]
    best = beam[0]
    return list(best[0]), -_lse2(*best[1])


def beam_search_decode(probs_btc, seq_len, blank, beam_width):
    """[B,T,C] probabilities -> (list of label lists, list of neg-log scores)."""
    res, sc = [], []
    with np.errstate(divide="ignore"):
        # the reference takes np.log in the dtype of `probs` (float32 as returned by
        # sess.run) and then accumulates in Python floats (beam_search_decoder.py:69)
        lp = np.log(np.asarray(probs_btc)).astype(np.float64)
    for b in range(len(seq_len)):
        h, s = beam_search_decode_single(lp[b, :int(seq_len[b])], blank, beam_width)
        res.append(h)
        sc.append(s)
    return res, sc


def edit_distance(hyp, ref):
    """Levenshtein distance (``tf.edit_distance`` semantics, ctc.py:391)."""
    n, m = len(hyp), len(ref)
    d = list(range(m + 1))
    for i in range(1, n + 1):
        prev, d[0] = d[0], i
        for j in range(1, m + 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (hyp[i - 1] != ref[j - 1]))
            prev = cur
    return d[m]


def label_error_rate(hyps, refs):
    """mean_b(edit_distance/len(ref))  (``compute_ler``, ctc.py:382-398)."""
    return float(np.mean([edit_distance(h, r) / len(r) for h, r in zip(hyps, refs)]))
