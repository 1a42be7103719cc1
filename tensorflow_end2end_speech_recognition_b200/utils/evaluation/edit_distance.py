"""PER / CER / WER, same names and call shapes as the reference's
``utils/evaluation/edit_distance.py:15-109`` -- without python-Levenshtein or TensorFlow: every distance is the
device Levenshtein kernel ``b2_edit_distance`` (``csrc/decode.cu``), one launch per call (batched forms below
take whole evaluation sets in one launch).

    compute_edit_distance(session, labels_true_st, labels_pred_st)   :15-33  (tf.edit_distance, normalize=True)
    compute_per(ref, hyp, normalize=True)                            :36-57  (Levenshtein over phone tokens)
    compute_cer(str_pred, str_true, normalize=True)                  :60-73  (Levenshtein over characters)
    compute_wer(ref, hyp, normalize=True)                            :76-109 (Levenshtein over words)
    wer_align(ref, hyp) -> (substitute, insert, delete)              :112-   (alignment counts)
"""
import numpy as np


def _device():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("edit distances are computed on the GPU (no CPU fallback)")
    return "cuda:%d" % torch.cuda.current_device()


def _tokens_to_ids(*seqs):
    """map arbitrary hashable tokens (phones, characters, words) to int32 ids, shared across the sequences"""
    table = {}
    out = []
    for s in seqs:
        out.append([table.setdefault(t, len(table)) for t in s])
    return out


def batch_edit_distance(hyps, refs, normalize=True, device=None):
    """hyps / refs: lists of token sequences (any hashable tokens) -> float64 array of (normalised) distances.
    One ``b2_edit_distance`` launch for the whole list."""
    from ... import ops
    assert len(hyps) == len(refs)
    if not len(hyps):
        return np.zeros(0)
    table = {}
    enc = lambda s: [table.setdefault(t, len(table)) for t in s]
    h = [enc(s) for s in hyps]
    r = [enc(s) for s in refs]
    d = ops.edit_distance(h, r, device or _device()).astype(np.float64)
    if normalize:
        with np.errstate(divide="ignore", invalid="ignore"):
            d = d / np.asarray([len(x) for x in refs], np.float64)
    return d


def compute_edit_distance(session, labels_true_st, labels_pred_st):
    """normalised edit distance per utterance of two sparse label batches (edit_distance.py:15-33).
    ``session`` is unused (kept for the call shape).  Note the reference swaps the two arguments when it builds its
    SparseTensors (prediction placeholder from the truth triple and vice versa, :24-27) and then calls
    ``tf.edit_distance(labels_pred_pl, labels_true_pl)``: the distance is symmetric, the normaliser is the length
    of what it names ``labels_true_pl`` = the PREDICTION -- reproduced here."""
    from ...utils.io.labels.sparsetensor import sparse_to_label_lists
    B = int(np.asarray(labels_true_st[2])[0])
    true_l = sparse_to_label_lists(labels_true_st, B)
    pred_l = sparse_to_label_lists(labels_pred_st, B)
    # hypothesis := truth triple, truth := prediction triple (the reference's swap)
    return batch_edit_distance(true_l, pred_l, normalize=True).astype(np.float32)


def compute_per(ref, hyp, normalize=True):
    """Phone Error Rate (edit_distance.py:36-57)."""
    d = float(batch_edit_distance([list(hyp)], [list(ref)], normalize=False)[0])
    return d / len(ref) if normalize else d


def compute_cer(str_pred, str_true, normalize=True):
    """Character Error Rate (edit_distance.py:60-73)."""
    d = float(batch_edit_distance([list(str_pred)], [list(str_true)], normalize=False)[0])
    return d / len(list(str_true)) if normalize else d


def compute_wer(ref, hyp, normalize=True):
    """Word Error Rate (edit_distance.py:76-109)."""
    d = float(batch_edit_distance([list(hyp)], [list(ref)], normalize=False)[0])
    return d / len(ref) if normalize else d


def wer_align(ref, hyp, verbose=False):
    """(substitute, insert, delete) counts of one optimal alignment, chosen with the reference's tie order
    (match, insertion, substitution, deletion; edit_distance.py:147-166).  The backtrace needs the full DP table,
    so this evaluation-report helper runs on the host (numpy); the total equals ``compute_wer(normalize=False)``."""
    n, m = len(ref), len(hyp)
    d = np.zeros((n + 1, m + 1), np.int32)
    d[0, :] = np.arange(m + 1)
    d[:, 0] = np.arange(n + 1)
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            if ref[i - 1] == hyp[j - 1]:
                d[i, j] = d[i - 1, j - 1]
            else:
                d[i, j] = min(d[i - 1, j - 1], d[i, j - 1], d[i - 1, j]) + 1
    x, y = n, m
    sub = ins = dele = 0
    ops_rev = []
    while x > 0 or y > 0:
        if x > 0 and y > 0 and d[x, y] == d[x - 1, y - 1] and ref[x - 1] == hyp[y - 1]:
            ops_rev.append("e"); x -= 1; y -= 1
        elif y > 0 and d[x, y] == d[x, y - 1] + 1:
            ops_rev.append("i"); ins += 1; y -= 1
        elif x > 0 and y > 0 and d[x, y] == d[x - 1, y - 1] + 1:
            ops_rev.append("s"); sub += 1; x -= 1; y -= 1
        else:
            ops_rev.append("d"); dele += 1; x -= 1
    if verbose:
        print("ALIGN:", "".join(ops_rev[::-1]))
    return sub, ins, dele
