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


# --------------------------------------------------------------------------- TF-semantics beam search
class _TfBeamEntry(object):
    __slots__ = ("parent", "label", "children", "old", "new")

    def __init__(self, parent, label):
        self.parent, self.label, self.children = parent, label, None
        self.old = [NEG_INF, NEG_INF, NEG_INF]      # total, blank, label   (log domain)
        self.new = [NEG_INF, NEG_INF, NEG_INF]

    def active(self):
        return self.new[0] != NEG_INF


def tf_ctc_beam_search_single(logits_tc, blank, beam_width, merge_repeated=True, top_paths=1,
                              normalize=False):
    """``tf.nn.ctc_beam_search_decoder`` for one utterance, as the reference calls it at ``models/ctc/ctc.py:344-346``
    (``beam_width`` from the config, ``top_paths=1``, ``merge_repeated=True`` by default).

    TensorFlow is not vendored in the reference and cannot be installed here: this restates the algorithm of
    TF 1.x ``tensorflow/core/util/ctc/ctc_beam_search.h`` (``CTCBeamSearchDecoder::Step`` / ``TopPaths``,
    ``ctc_beam_entry.h::LabelSeq``) from its published source -- PARITY UNPINNED against TensorFlow itself:
      * per frame the input is the logit row minus its maximum (TF <= 1.8; ``normalize=True`` gives the
        log-softmax of later versions -- a per-frame constant, rankings are identical);
      * the beam is a prefix tree; every leaf keeps (P_total, P_blank, P_label) at t-1 and t;
      * existing leaves are updated first (label part from the parent only while the parent is still in the beam),
        then every leaf whose OLD total still beats the current bottom of the beam proposes its children, in
        descending old-probability order, labels ascending; a child enters iff its total beats the bottom;
      * ``blank`` must be the last class (children are labels 0..C-2);
      * the returned path walks leaf -> root and, with ``merge_repeated``, drops a label equal to the one emitted
        right after it (so "a a b" comes out as "a b" -- TF's documented quirk).
    Ties between equal totals are broken by the position in the candidate list (TF leaves them to its heap).
    Returns ([label lists, best first], [log scores])."""
    T, C = logits_tc.shape
    assert blank == C - 1, "TF's decoder assumes the blank is the last class"
    root = _TfBeamEntry(None, -1)
    root.new = [0.0, 0.0, NEG_INF]
    leaves = [root]
    for t in range(T):
        row = np.asarray(logits_tc[t], np.float64)
        x = row - row.max()
        if normalize:
            x = x - math.log(np.exp(x).sum())
        branches = sorted(leaves, key=lambda e: -e.new[0])          # Extract(): descending new probability
        leaves = []
        for b in branches:
            b.old = list(b.new)
        for b in branches:
            if b.parent is not None:
                if b.parent.active():
                    prev = b.parent.old[1] if b.label == b.parent.label else b.parent.old[0]
                    b.new[2] = _lse2(b.new[2], prev)
                b.new[2] += x[b.label]
            b.new[1] = b.old[0] + x[blank]
            b.new[0] = _lse2(b.new[1], b.new[2])
            leaves.append(b)

        def bottom():
            return min(leaves, key=lambda e: e.new[0])

        def is_candidate(total):
            return total > NEG_INF and (len(leaves) < beam_width or total > bottom().new[0])

        def push(e):
            if len(leaves) < beam_width:
                leaves.append(e)
            else:
                bt = bottom()
                if e.new[0] > bt.new[0]:
                    bt.new = [NEG_INF, NEG_INF, NEG_INF]
                    leaves.remove(bt)
                    leaves.append(e)
        # (existing leaves beyond the width cannot happen: |branches| <= beam_width)
        for b in branches:
            if not is_candidate(b.old[0]):
                continue
            if b.children is None:
                b.children = [_TfBeamEntry(b, c) for c in range(C - 1)]
            for c in b.children:
                if c.active():
                    continue
                prev = b.old[1] if c.label == b.label else b.old[0]
                c.new = [x[c.label] + prev, NEG_INF, x[c.label] + prev]
                if is_candidate(c.new[0]):
                    push(c)
                else:
                    c.old = [NEG_INF, NEG_INF, NEG_INF]
                    c.new = [NEG_INF, NEG_INF, NEG_INF]
    best = sorted(leaves, key=lambda e: -e.new[0])[:top_paths]
    paths, scores = [], []
    for e in best:
        labels, prev_label, c = [], -1, e
        while c.parent is not None:
            if not merge_repeated or c.label != prev_label:
                labels.append(c.label)
            prev_label = c.label
            c = c.parent
        paths.append(labels[::-1])
        scores.append(e.new[0])
    return paths, scores


def tf_ctc_beam_search(logits_tbc, seq_len, blank, beam_width, merge_repeated=True):
    """batched form over time-major logits [T,B,C] -> list of best label lists"""
    out = []
    for b in range(logits_tbc.shape[1]):
        paths, _ = tf_ctc_beam_search_single(np.asarray(logits_tbc[:int(seq_len[b]), b]), blank, beam_width,
                                             merge_repeated)
        out.append(paths[0])
    return out
