"""Attention layer (energy -> masked softmax -> context), CPU restatement.  TEST INFRASTRUCTURE.

Follows ``models/attention/decoders/attention_layer.py``: ``__call__`` :45-113 (mask with
``tf.float32.min`` :84-85, ``*= sharpening_factor`` :89, softmax or sigmoid/sum :92-99, context
:106-111) and ``_compute_attention_score`` :115-347 for the implemented types
``dot_product`` (:162-171), ``bahdanau_content`` (:173-189), ``hybrid`` (:191-229, conv filter
[200,1,10]), ``location`` (:231-265, filter [201,1,10]), ``luong_dot`` (:271-288),
``luong_general`` (:290-312), ``luong_concat`` (:314-345).  ``tf.nn.conv1d(..., padding='SAME')``
is a cross-correlation with pad_left = (k-1)//2 (TF-upstream).  The reference feeds
all-zero previous weights to location/hybrid (A.7.1); callers choose what to pass.

params (numpy arrays, TF variable names):
  W_query/weights [Dq,A]            (types that project the query)
  W_keys/weights [E,A] (+ W_keys/biases [A] unless dot_product) ; luong_general: [E,Dq] no bias
  filter [k,1,10], W_filter/weights [10,A], W_filter/biases [A]   (hybrid, location)
  W_concat/weights [E+Dq, A]        (luong_concat)
  v_a [A]
"""
import numpy as np

FLOAT32_MIN = float(np.finfo(np.float32).min)
ATTENTION_TYPE = ["bahdanau_content", "location", "hybrid", "dot_product",
                  "luong_dot", "luong_general", "luong_concat"]


def conv1d_same(x_bt, filt_k10):
    """x [B,T] -> [B,T,10]; cross-correlation, SAME padding (left = (k-1)//2)."""
    B, T = x_bt.shape
    k = filt_k10.shape[0]
    pl = (k - 1) // 2
    xp = np.zeros((B, T + k - 1), dtype=np.float64)
    xp[:, pl:pl + T] = x_bt
    out = np.zeros((B, T, filt_k10.shape[1]))
    for j in range(k):
        out += xp[:, j:j + T, None] * filt_k10[j][None, None, :]
    return out


def attention_energy(attention_type, enc, query, prev_alpha, p):
    """enc [B,T,E], query [B,Dq], prev_alpha [B,T] -> energy [B,T] (float64)."""
    enc = np.asarray(enc, np.float64)
    query = np.asarray(query, np.float64)
    g = lambda k: np.asarray(p[k], np.float64)
    if attention_type not in ATTENTION_TYPE:
        raise ValueError("attention type should be one of [%s], you provided %s." %
                         (", ".join(ATTENTION_TYPE), attention_type))
    if attention_type in ("bahdanau_content", "location", "hybrid", "dot_product"):
        wq = query @ g("W_query/weights")                       # [B,A]
        wk = enc @ g("W_keys/weights")                          # [B,T,A]
        if attention_type != "dot_product":
            wk = wk + g("W_keys/biases")
        if attention_type == "dot_product":
            return np.einsum("bta,ba->bt", wk, wq)
        if attention_type == "bahdanau_content":
            return np.sum(g("v_a") * np.tanh(wk + wq[:, None, :]), axis=2)
        f = conv1d_same(np.asarray(prev_alpha, np.float64), g("filter")[:, 0, :])
        wf = f @ g("W_filter/weights") + g("W_filter/biases")
        if attention_type == "hybrid":
            return np.sum(g("v_a") * np.tanh(wk + wq[:, None, :] + wf), axis=2)
        return np.sum(g("v_a") * np.tanh(wq[:, None, :] + wf), axis=2)     # location
    if attention_type == "luong_dot":
        if enc.shape[-1] != query.shape[-1]:
            raise ValueError("encoder_num_units and decoder_num_units must be the same size.")
        return np.einsum("bte,be->bt", enc, query)
    if attention_type == "luong_general":
        return np.einsum("btd,bd->bt", enc @ g("W_keys/weights"), query)
    T = enc.shape[1]
    cat = np.concatenate([enc, np.repeat(query[:, None, :], T, axis=1)], axis=2)
    return np.sum(g("v_a") * np.tanh(cat @ g("W_concat/weights")), axis=2)   # luong_concat


def attention_step(attention_type, enc, query, enc_len, prev_alpha, p, sharpening_factor=1.0,
                   sigmoid_smoothing=False):
    """-> (attention_weights [B,T], context [B,E]) as float64."""
    enc = np.asarray(enc, np.float64)
    B, T, _ = enc.shape
    e = attention_energy(attention_type, enc, query, prev_alpha, p)
    mask = (np.arange(T)[None, :] < np.asarray(enc_len)[:, None]).astype(np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        e = e * mask + (1.0 - mask) * FLOAT32_MIN
        e = e * sharpening_factor
        if sigmoid_smoothing:
            s = 1.0 / (1.0 + np.exp(-np.clip(e, -745, 745)))
            s = np.where(mask > 0, s, 0.0)
            w = s / s.sum(-1, keepdims=True)
        else:
            m = e.max(-1, keepdims=True)
            ex = np.exp(e - m)
            w = ex / ex.sum(-1, keepdims=True)
    return w, np.einsum("bt,bte->be", w, enc)
