"""CTC loss + gradient, CPU restatement (numpy, float64).  TEST INFRASTRUCTURE.

Follows the reference call ``tf.nn.ctc_loss(labels, logits, seq_len,
preprocess_collapse_repeated=False, ctc_merge_repeated=True,
ignore_longer_outputs_than_inputs=True, time_major=True)`` at
``models/ctc/ctc.py:289-297`` (and ``models/attention/joint_ctc_attention.py:308-316``
with ``ignore_longer_outputs_than_inputs=False``).  The arithmetic itself lives
in TensorFlow 1.x (pinned ``tensorflow==1.2.0``, ``requirements.txt:11``), which
is absent; this restates the published algorithm (Graves 2006, as implemented by
TF's ``ctc_loss_calculator``): blank = C-1, extended label l' = [b,l1,b,...,lL,b],
softmax then log, alpha/beta in log space, gradient wrt the *unnormalised*
logits ``y - occupancy/p``.  TF-upstream semantics restated:

* frames ``t >= seq_len[b]`` get zero gradient;
* a label sequence that cannot be emitted (``L + repeats > seq_len``) has
  ``log p = -inf`` -> loss ``+inf`` and gradient = softmax (TF prints a warning);
* ``ignore_longer_outputs_than_inputs=True``: utterances with ``L > seq_len``
  are skipped: loss 0, gradient 0;  ``False``: ValueError (TF: InvalidArgument).
"""
import numpy as np

NEG_INF = -np.inf


def _lse(*xs):
    m = max(xs)
    if m == NEG_INF:
        return NEG_INF
    return m + np.log(sum(np.exp(x - m) for x in xs))


def log_softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=axis, keepdims=True))


def ctc_loss_single(logits_tc, labels, blank):
    """One utterance.  logits_tc: [T, C] float64; labels: 1-D ints (no blanks).

    Returns (neg_log_p, grad[T, C], alpha[T,S], beta[T,S]).  alpha and beta both
    *include* the emission at t, so  sum_s alpha_t(s) beta_t(s) / y_t(l'(s)) = p.
    """
    T, C = logits_tc.shape
    L = len(labels)
    S = 2 * L + 1
    lp = log_softmax(logits_tc.astype(np.float64))
    ext = np.full(S, blank, dtype=np.int64)
    ext[1::2] = labels
    alpha = np.full((T, S), NEG_INF)
    beta = np.full((T, S), NEG_INF)
    if T == 0:
        return (0.0 if L == 0 else np.inf), np.zeros_like(lp), alpha, beta
    alpha[0, 0] = lp[0, blank]
    if S > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        for s in range(S):
            a = alpha[t - 1, s]
            if s >= 1:
                a = _lse(a, alpha[t - 1, s - 1])
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                a = _lse(a, alpha[t - 1, s - 2])
            alpha[t, s] = a + lp[t, ext[s]]
    beta[T - 1, S - 1] = lp[T - 1, blank]
    if S > 1:
        beta[T - 1, S - 2] = lp[T - 1, ext[S - 2]]
    for t in range(T - 2, -1, -1):
        for s in range(S):
            b = beta[t + 1, s]
            if s + 1 < S:
                b = _lse(b, beta[t + 1, s + 1])
            if s + 2 < S and ext[s + 2] != blank and ext[s + 2] != ext[s]:
                b = _lse(b, beta[t + 1, s + 2])
            beta[t, s] = b + lp[t, ext[s]]
    logp = alpha[T - 1, S - 1]
    if S > 1:
        logp = _lse(logp, alpha[T - 1, S - 2])
    y = np.exp(lp)
    if logp == NEG_INF:
        return np.inf, y, alpha, beta
    occ = np.zeros((T, C))
    for t in range(T):
        for s in range(S):
            v = alpha[t, s] + beta[t, s]
            if v > NEG_INF:
                occ[t, ext[s]] += np.exp(v - lp[t, ext[s]] - logp)
    return -logp, y - occ, alpha, beta


def ctc_loss(logits, labels, seq_len, blank=None, ignore_longer_outputs_than_inputs=True):
    """Batch CTC, TF layout.

    logits : [T, B, C] (time-major, unnormalised)   (ctc.py:226-236 builds this)
    labels : list of B int sequences (dense form of the SparseTensor, sparsetensor.py:12-39)
    seq_len: [B] ints
    Returns (loss[B] float64, grad[T,B,C] float64 = d sum_b(loss_b) / d logits).
    """
    logits = np.asarray(logits, dtype=np.float64)
    T, B, C = logits.shape
    if blank is None:
        blank = C - 1
    loss = np.zeros(B)
    grad = np.zeros_like(logits)
    for b in range(B):
        tb = int(seq_len[b])
        lab = np.asarray(labels[b], dtype=np.int64)
        if len(lab) > tb:
            if ignore_longer_outputs_than_inputs:
                continue
            raise ValueError("Not enough time for target transition sequence "
                             "(required: %d, available: %d)" % (len(lab), tb))
        nll, g, _, _ = ctc_loss_single(logits[:tb, b], lab, blank)
        loss[b] = nll
        grad[:tb, b] = g
    return loss, grad


def ctc_brute_force(logits_tc, labels, blank):
    """-log p by enumerating every alignment; only for T<=7, C<=4 (tests)."""
    import itertools
    T, C = logits_tc.shape
    y = np.exp(log_softmax(logits_tc.astype(np.float64)))
    p = 0.0
    for path in itertools.product(range(C), repeat=T):
        col = [k for k, _ in itertools.groupby(path)]
        col = [k for k in col if k != blank]
        if col == list(labels):
            p += np.prod([y[t, path[t]] for t in range(T)])
    return -np.log(p) if p > 0 else np.inf


def _lse_vec(*xs):
    m = np.maximum.reduce(xs)
    safe = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore"):
        out = safe + np.log(sum(np.exp(x - safe) for x in xs))
    return np.where(np.isfinite(m), out, NEG_INF)


def ctc_loss_single_fast(logits_tc, labels, blank):
    """Same as ``ctc_loss_single`` but vectorised over the lattice axis s
    (usable at T=1000, S=401).  Returns (neg_log_p, grad[T, C])."""
    T, C = logits_tc.shape
    L = len(labels)
    S = 2 * L + 1
    lp = log_softmax(logits_tc.astype(np.float64))
    ext = np.full(S, blank, dtype=np.int64)
    ext[1::2] = labels
    skip = np.zeros(S, dtype=bool)           # may come from s-2
    skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    lpe = lp[:, ext]                          # [T, S]
    alpha = np.full((T, S), NEG_INF)
    beta = np.full((T, S), NEG_INF)
    alpha[0, 0] = lpe[0, 0]
    if S > 1:
        alpha[0, 1] = lpe[0, 1]
    pad1 = np.array([NEG_INF])
    pad2 = np.array([NEG_INF, NEG_INF])
    for t in range(1, T):
        a0 = alpha[t - 1]
        a1 = np.concatenate([pad1, a0[:-1]])
        a2 = np.where(skip, np.concatenate([pad2, a0[:-2]]), NEG_INF)
        alpha[t] = _lse_vec(a0, a1, a2) + lpe[t]
    beta[T - 1, S - 1] = lpe[T - 1, S - 1]
    if S > 1:
        beta[T - 1, S - 2] = lpe[T - 1, S - 2]
    skipb = np.zeros(S, dtype=bool)           # may go to s+2
    skipb[:-2] = skip[2:]
    for t in range(T - 2, -1, -1):
        b0 = beta[t + 1]
        b1 = np.concatenate([b0[1:], pad1])
        b2 = np.where(skipb, np.concatenate([b0[2:], pad2]), NEG_INF)
        beta[t] = _lse_vec(b0, b1, b2) + lpe[t]
    logp = alpha[T - 1, S - 1]
    if S > 1:
        logp = _lse(logp, alpha[T - 1, S - 2])
    y = np.exp(lp)
    if logp == NEG_INF:
        return np.inf, y
    with np.errstate(under="ignore"):
        w = np.exp(alpha + beta - lpe - logp)   # [T, S]; exp(-inf) = 0
    occ = np.zeros((T, C))
    for s in range(S):
        occ[:, ext[s]] += w[:, s]
    return -logp, y - occ


def ctc_loss_fast(logits, labels, seq_len, blank=None, ignore_longer_outputs_than_inputs=True):
    """Vectorised twin of ``ctc_loss`` (identical contract)."""
    logits = np.asarray(logits, dtype=np.float64)
    T, B, C = logits.shape
    if blank is None:
        blank = C - 1
    loss = np.zeros(B)
    grad = np.zeros_like(logits)
    for b in range(B):
        tb = int(seq_len[b])
        lab = np.asarray(labels[b], dtype=np.int64)
        if len(lab) > tb:
            if ignore_longer_outputs_than_inputs:
                continue
            raise ValueError("Not enough time for target transition sequence "
                             "(required: %d, available: %d)" % (len(lab), tb))
        if tb == 0:
            loss[b] = 0.0 if len(lab) == 0 else np.inf
            continue
        nll, g = ctc_loss_single_fast(logits[:tb, b], lab, blank)
        loss[b] = nll
        grad[:tb, b] = g
    return loss, grad
