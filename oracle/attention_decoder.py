"""Attention decoder loop (bridge, input-feeding LSTM step, helpers), CPU restatement.
TEST INFRASTRUCTURE -- float64 numpy, small sizes only.

Follows
  ``models/attention/bridge.py`` InitialStateBridge._create: concat of the flattened encoder
      final state ((c_fw, h_fw), (c_bw, h_bw)) -> fully_connected (identity, bias) -> split into
      the decoder cell state (c0, h0);
  ``models/attention/decoders/attention_decoder.py``
      initialize :143-168 (first input = [emb(<SOS>) ; zero context], attention weights = 0),
      _compute_output :170-211 (attention -> tanh(W_av [cell_out ; ctx]) no bias -> logits FC),
      step :256-295 (cell -> _compute_output -> helper.sample -> next input = [emb ; ctx]);
  ``models/attention/decoders/dynamic_decoder.py:148-212`` (loop until all finished or
      ``maximum_iterations``; impute_finished: zero outputs and copy state through for finished
      sequences);
  ``models/attention/attention_seq2seq.py:352-363`` (LSTMBlockCell, forget_bias 1),
      :432-446 (TrainingHelper on labels[:, :-1], length labels_seq_len - 1),
      :486-494 (GreedyEmbeddingHelper(<SOS>, <EOS>)).

TF-upstream facts restated (tf.contrib.seq2seq 1.x, not vendored): GreedyEmbeddingHelper --
sample = argmax(logits) (first max), finished = (sample == end_token), next input =
emb(sample); TrainingHelper -- sample = argmax, finished = (time + 1 >= sequence_length), next
input = inputs[:, time + 1] (zeros once every sequence is finished).

The decoder object's ``self.attention_weights`` is read inside a ``tf.while_loop`` body that is
traced once, so every step is fed the all-zero weights of ``initialize``;
``feed_previous_attention=True`` gives the intended recurrence instead.

params (numpy, TF variable names):
  bridge/weights [4*H_enc, 2*Hd], bridge/biases [2*Hd]
  W_embedding [num_classes, emb]
  cell: kernel [emb + E + Hd, 4*Hd], bias [4*Hd], w_{i,f,o}_diag [Hd] (optional)
  attentional_vector/weights [Hd + E, Hd]
  output_layer/weights [Hd, num_classes], output_layer/biases [num_classes]
  attention: dict for oracle.attention.attention_step
"""
import numpy as np

from .attention import attention_step


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def bridge_initial_state(final_state, p):
    """final_state = ((c_fw, h_fw), (c_bw, h_bw)) each [B, H_enc] -> (c0, h0) [B, Hd]."""
    flat = np.concatenate([np.asarray(s, np.float64) for pair in final_state for s in pair], axis=1)
    out = flat @ np.asarray(p["bridge/weights"], np.float64) + np.asarray(p["bridge/biases"], np.float64)
    Hd = out.shape[1] // 2
    return out[:, :Hd], out[:, Hd:]


def cell_step(x, c_prev, h_prev, cell, forget_bias=1.0, cell_clip=None):
    g = lambda k: np.asarray(cell[k], np.float64)
    z = np.concatenate([x, h_prev], axis=1) @ g("kernel") + g("bias")
    i, gg, f, o = np.split(z, 4, axis=1)
    if "w_i_diag" in cell:
        c = _sig(f + forget_bias + g("w_f_diag") * c_prev) * c_prev + _sig(i + g("w_i_diag") * c_prev) * np.tanh(gg)
    else:
        c = _sig(f + forget_bias) * c_prev + _sig(i) * np.tanh(gg)
    if cell_clip:
        c = np.clip(c, -cell_clip, cell_clip)
    if "w_o_diag" in cell:
        h = _sig(o + g("w_o_diag") * c) * np.tanh(c)
    else:
        h = _sig(o) * np.tanh(c)
    return c, h


def decode(p, attention_type, enc, enc_len, initial_state, sos=None, eos=None, max_decode_length=None,
           labels=None, labels_seq_len=None, sharpening_factor=1.0, sigmoid_smoothing=False,
           feed_previous_attention=False, cell_clip=None, forced_ids=None):
    """Greedy (labels is None) or teacher-forced decoding.

    Returns dict of batch-major arrays: logits [B,L,C], predicted_ids [B,L], decoder_output
    [B,L,Hd], attention_weights [B,L,T], context_vector [B,L,E], plus final (c, h).
    ``forced_ids`` [B,L] (greedy mode only) overrides the sampled id fed back at each step, so a
    GPU run whose arg-max differs by a float tie can still be followed step by step.
    """
    enc = np.asarray(enc, np.float64)
    B, T, E = enc.shape
    emb = np.asarray(p["W_embedding"], np.float64)
    w_av = np.asarray(p["attentional_vector/weights"], np.float64)
    w_o = np.asarray(p["output_layer/weights"], np.float64)
    b_o = np.asarray(p["output_layer/biases"], np.float64)
    C = w_o.shape[1]
    c, h = [np.array(s, np.float64) for s in initial_state]
    Hd = h.shape[1]
    teacher = labels is not None
    if teacher:
        labels = np.asarray(labels)
        dec_in = emb[labels[:, :-1]]                            # [B, L, emb]
        seq = np.asarray(labels_seq_len) - 1
        finished = seq <= 0
        x_emb = dec_in[:, 0]
        max_iter = None
    else:
        finished = np.zeros(B, bool)
        x_emb = emb[np.full(B, sos)]
        max_iter = max_decode_length
        if max_iter is not None and max_iter <= 0:
            finished[:] = True
    ctx = np.zeros((B, E))
    alpha_state = np.zeros((B, T))
    outs = {k: [] for k in ("logits", "predicted_ids", "decoder_output", "attention_weights", "context_vector")}
    time = 0
    while not finished.all():
        x = np.concatenate([x_emb, ctx], axis=1)
        c_new, h_new = cell_step(x, c, h, p["cell"], 1.0, cell_clip)
        alpha, ctx_new = attention_step(attention_type, enc, h_new, enc_len,
                                        alpha_state if feed_previous_attention else np.zeros((B, T)),
                                        p["attention"], sharpening_factor, sigmoid_smoothing)
        av = np.tanh(np.concatenate([h_new, ctx_new], axis=1) @ w_av)
        logits = av @ w_o + b_o
        ids = logits.argmax(1)
        if teacher:
            step_fin = (time + 1) >= seq
            nxt = dec_in[:, time + 1] if (time + 1) < dec_in.shape[1] and not step_fin.all() else np.zeros_like(x_emb)
        else:
            feed = ids if forced_ids is None else np.asarray(forced_ids)[:, time]
            step_fin = feed == eos
            nxt = emb[feed]
        keep = ~finished
        m = keep[:, None].astype(np.float64)
        outs["logits"].append(logits * m)
        outs["predicted_ids"].append(np.where(keep, ids, 0))
        outs["decoder_output"].append(av * m)
        outs["attention_weights"].append(alpha * m)
        outs["context_vector"].append(ctx_new * m)
        c = np.where(keep[:, None], c_new, c)
        h = np.where(keep[:, None], h_new, h)
        alpha_state = alpha
        ctx = ctx_new                         # next input uses the un-imputed context (:232-235)
        x_emb = nxt
        finished = finished | step_fin
        time += 1
        if max_iter is not None and time >= max_iter:
            finished[:] = True
    res = {k: (np.stack(v, axis=1) if v else np.zeros((B, 0))) for k, v in outs.items()}
    res["final_state"] = (c, h)
    return res


def beam_search_decode(p, attention_type, enc, enc_len, initial_state, sos, eos, beam_width,
                       length_penalty_weight=0.6, max_decode_length=100, sharpening_factor=1.0,
                       sigmoid_smoothing=False, feed_previous_attention=False, cell_clip=None):
    """Beam search for ONE utterance (enc [T,E], initial_state (c0,h0) [Hd]) following
    ``beam_search/beam_search_decoder.py:234-332`` (beam_search_step), ``util.py:38-95``
    (mask_probs with float32.min, normalize_score -- disabled for weight None / 1 --, top-k with the
    lower flat index first among equals) and ``util.py:14-26`` (gather_tree).  At time 0 only beam 0 is
    expanded.  The loop ends when every beam is finished or at ``max_decode_length``.
    Returns dict(ids [W,L], lengths [W], log_probs [W], scores [W], min_margin): ``min_margin`` is the
    smallest score gap between the last kept and the first rejected candidate over all steps (how
    close the search came to a tie)."""
    from .attention import FLOAT32_MIN
    W = beam_width
    enc = np.asarray(enc, np.float64)[None].repeat(W, axis=0)            # the beam is the batch
    T, E = enc.shape[1], enc.shape[2]
    lens = np.full(W, int(enc_len))
    emb = np.asarray(p["W_embedding"], np.float64)
    w_av = np.asarray(p["attentional_vector/weights"], np.float64)
    w_o = np.asarray(p["output_layer/weights"], np.float64)
    b_o = np.asarray(p["output_layer/biases"], np.float64)
    C = w_o.shape[1]
    c = np.tile(np.asarray(initial_state[0], np.float64)[None], (W, 1))
    h = np.tile(np.asarray(initial_state[1], np.float64)[None], (W, 1))
    x_emb = emb[np.full(W, sos)]
    ctx = np.zeros((W, E))
    alpha_prev = np.zeros((W, T))
    log_probs, finished, lengths = np.zeros(W), np.zeros(W, bool), np.zeros(W, int)
    hist_w, hist_p = [], []
    min_margin = np.inf
    scores_sel = np.zeros(W)
    for time in range(max_decode_length):
        x = np.concatenate([x_emb, ctx], axis=1)
        c_new, h_new = cell_step(x, c, h, p["cell"], 1.0, cell_clip)
        alpha, ctx_new = attention_step(attention_type, enc, h_new, lens,
                                        alpha_prev if feed_previous_attention else np.zeros((W, T)),
                                        p["attention"], sharpening_factor, sigmoid_smoothing)
        av = np.tanh(np.concatenate([h_new, ctx_new], axis=1) @ w_av)
        logits = av @ w_o + b_o
        m = logits.max(1, keepdims=True)
        probs = logits - (m + np.log(np.exp(logits - m).sum(1, keepdims=True)))
        fin_row = np.full(C, FLOAT32_MIN)
        fin_row[eos] = 0.0
        probs = np.where(finished[:, None], fin_row[None], probs)
        total = log_probs[:, None] + probs
        add = np.ones((W, C), int)
        add[:, eos] = 0
        new_len = lengths[:, None] + (~finished)[:, None].astype(int) * add
        if length_penalty_weight is None or length_penalty_weight == 1:
            scores = total
        else:
            scores = total / ((5.0 + new_len) ** length_penalty_weight / 6.0 ** length_penalty_weight)
        flat = scores.reshape(-1) if time > 0 else scores[0]
        order = np.argsort(-flat, kind="stable")                          # ties: lower index first
        idx = order[:W]
        if len(order) > W:
            min_margin = min(min_margin, float(flat[order[W - 1]] - flat[order[W]]))
        min_margin = min(min_margin, float(np.min(-np.diff(flat[idx]))) if W > 1 else np.inf)
        scores_sel = flat[idx]
        log_probs = total.reshape(-1)[idx]
        words, parents = idx % C, idx // C
        nf = finished[parents] | (words == eos)
        lengths = lengths[parents] + (~nf).astype(int) * (words != eos).astype(int)
        finished = nf
        hist_w.append(words)
        hist_p.append(parents)
        c, h, ctx, alpha_prev = c_new[parents], h_new[parents], ctx_new[parents], alpha[parents]
        x_emb = emb[words]
        if finished.all():
            break
    L = len(hist_w)
    ids = np.zeros((W, L), int)
    for w in range(W):
        cur = w
        for t in range(L - 1, -1, -1):
            ids[w, t] = hist_w[t][cur]
            cur = hist_p[t][cur]
    return {"ids": ids, "lengths": lengths, "log_probs": log_probs, "scores": scores_sel,
            "min_margin": min_margin}
