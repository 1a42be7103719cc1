"""Stacked bidirectional LSTM encoder, CPU restatement.  TEST INFRASTRUCTURE.

Follows ``models/encoders/core/blstm.py:258-332`` (``lstmblockcell``; the
``lstmcell`` / ``basiclstmcell`` variants at :187-255 / :124-184 are the same
stack with flags) and the only in-tree statement of the cell equations,
``models/recurrent/layers/lstm.py:142-183``:

    [i, g, f, o] = split([x, h_prev] . W + b, 4)                      (:143-144)
    c = sig(f + forget_bias + w_f*c_prev)*c_prev + sig(i + w_i*c_prev)*tanh(g)   (:157-158)
    c = clip(c, +-cell_clip)                                          (:163-164)
    h = sig(o + w_o*c) * tanh(c)          (new c)                     (:166-167)
    h = h . W_proj                        (LSTMCell only)             (:171-176)

TF-upstream facts restated (not vendored in the reference):
``tf.contrib.rnn.LSTMBlockCell`` uses the same equations with gate column order
i, ci(=g), f, o, weight rows = [x; h], ``forget_bias`` added at run time;
``tf.nn.bidirectional_dynamic_rnn(sequence_length=len)`` starts from zero state,
emits zeros and copies the state through for ``t >= len_b``, and the backward
direction runs on ``reverse_sequence(x, len)`` (so it starts at ``t = len_b-1``).
``DropoutWrapper(output_keep_prob)`` scales only the emitted output.

Parameter container (one dict per layer and direction, TF variable names):
    kernel  [(D_in + H_out_prev) , 4H]   bias [4H]
    w_i_diag, w_f_diag, w_o_diag [H]     (peephole cells)
    projection [H, P]                     (LSTMCell with num_proj)
"""
import numpy as np
import torch


def init_blstm_params(input_size, num_units, num_layers, parameter_init=0.1,
                      use_peephole=True, num_proj=None, seed=0, dtype=np.float32):
    """U(-parameter_init, parameter_init) kernels/peepholes, zero biases
    (blstm.py:79-80, 283-284; TF default zero bias)."""
    rng = np.random.RandomState(seed)
    layers = []
    d_in = input_size
    out = num_proj if num_proj else num_units
    for _ in range(num_layers):
        layer = {}
        for d in ("fw", "bw"):
            p = {
                "kernel": rng.uniform(-parameter_init, parameter_init,
                                      (d_in + out, 4 * num_units)).astype(dtype),
                "bias": np.zeros(4 * num_units, dtype=dtype),
            }
            if use_peephole:
                for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
                    p[k] = rng.uniform(-parameter_init, parameter_init, num_units).astype(dtype)
            if num_proj:
                p["projection"] = rng.uniform(-parameter_init, parameter_init,
                                              (num_units, num_proj)).astype(dtype)
            layer[d] = p
        layers.append(layer)
        d_in = 2 * out
    return layers


def lstm_cell_step(x, h_prev, c_prev, p, forget_bias=1.0, cell_clip=None):
    """One cell step on torch tensors.  x [B,D], h_prev [B,Hout], c_prev [B,H]."""
    z = torch.cat([x, h_prev], dim=1) @ p["kernel"] + p["bias"]
    i, g, f, o = torch.chunk(z, 4, dim=1)
    if "w_i_diag" in p:
        c = torch.sigmoid(f + forget_bias + p["w_f_diag"] * c_prev) * c_prev + \
            torch.sigmoid(i + p["w_i_diag"] * c_prev) * torch.tanh(g)
    else:
        c = torch.sigmoid(f + forget_bias) * c_prev + torch.sigmoid(i) * torch.tanh(g)
    if cell_clip is not None:
        c = torch.clamp(c, -cell_clip, cell_clip)
    if "w_o_diag" in p:
        h = torch.sigmoid(o + p["w_o_diag"] * c) * torch.tanh(c)
    else:
        h = torch.sigmoid(o) * torch.tanh(c)
    if "projection" in p:
        h = h @ p["projection"]
    return h, c


def _run_direction(x_tbd, seq_len, p, reverse, forget_bias, cell_clip):
    """dynamic_rnn with sequence_length on a time-major input.  Returns
    (outputs [T,B,Hout] zero past len, c_final [B,H], h_final [B,Hout])."""
    T, B, _ = x_tbd.shape
    H = p["bias"].shape[0] // 4
    Hout = p["projection"].shape[1] if "projection" in p else H
    h = x_tbd.new_zeros(B, Hout)
    c = x_tbd.new_zeros(B, H)
    outs = [None] * T
    lens = torch.as_tensor(np.asarray(seq_len), dtype=torch.long)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        # reverse_sequence semantics: step t is active iff t < len_b (for both
        # directions; the backward one simply visits t in decreasing order and
        # keeps its zero state until it reaches t = len_b - 1).
        m = (t < lens).to(x_tbd.dtype).unsqueeze(1)
        h_new, c_new = lstm_cell_step(x_tbd[t], h, c, p, forget_bias, cell_clip)
        outs[t] = h_new * m
        h = h_new * m + h * (1 - m)
        c = c_new * m + c * (1 - m)
    return torch.stack(outs, 0), c, h


def blstm_forward(x_btd, seq_len, layers, keep_prob=1.0, dropout_masks=None,
                  forget_bias=1.0, cell_clip=None):
    """x_btd [B,T,D] torch tensor -> (outputs [T,B,2*Hout] time-major,
    final_state ((c_fw,h_fw),(c_bw,h_bw)) of the last layer).  (blstm.py:277-332)

    dropout_masks: optional list (per layer) of [T,B,2*Hout] {0,1} masks applied
    as ``out*mask/keep_prob`` to the emitted output only (DropoutWrapper).
    """
    x = x_btd.transpose(0, 1)                     # blstm.py:279
    final = None
    for li, layer in enumerate(layers):
        pf = {k: torch.as_tensor(v) if not torch.is_tensor(v) else v for k, v in layer["fw"].items()}
        pb = {k: torch.as_tensor(v) if not torch.is_tensor(v) else v for k, v in layer["bw"].items()}
        of, cf, hf = _run_direction(x, seq_len, pf, False, forget_bias, cell_clip)
        ob, cb, hb = _run_direction(x, seq_len, pb, True, forget_bias, cell_clip)
        x = torch.cat([of, ob], dim=2)            # blstm.py:323
        if dropout_masks is not None and keep_prob < 1.0:
            x = x * dropout_masks[li] / keep_prob
        final = ((cf, hf), (cb, hb))
    return x, final


def lstm_forward(x_btd, seq_len, layers, keep_prob=1.0, dropout_masks=None, forget_bias=1.0, cell_clip=None):
    """Unidirectional stack (models/encoders/core/lstm.py:120-166: MultiRNNCell of DropoutWrapper(LSTM cells) under
    dynamic_rnn with sequence_length).  layers: list of parameter dicts.  -> (outputs [T,B,H] time-major,
    ((c, h) per layer))."""
    x = x_btd.transpose(0, 1)
    states = []
    for li, p in enumerate(layers):
        p = {k: torch.as_tensor(v) if not torch.is_tensor(v) else v for k, v in p.items()}
        x, c, h = _run_direction(x, seq_len, p, False, forget_bias, cell_clip)
        if dropout_masks is not None and keep_prob < 1.0:
            x = x * dropout_masks[li] / keep_prob
        states.append((c, h))
    return x, tuple(states)


def blstm_forward_numpy(x_btd, seq_len, layers, forget_bias=1.0, cell_clip=None):
    """Literal numpy loop (float64), independent of the torch form; used only to
    cross-check ``blstm_forward`` on tiny shapes."""
    def sig(v):
        return 1.0 / (1.0 + np.exp(-v))
    x = np.transpose(np.asarray(x_btd, np.float64), (1, 0, 2))
    T, B, _ = x.shape
    for layer in layers:
        outs = []
        for d, rev in (("fw", False), ("bw", True)):
            p = {k: np.asarray(v, np.float64) for k, v in layer[d].items()}
            H = p["bias"].shape[0] // 4
            out = np.zeros((T, B, H))
            for b in range(B):
                h = np.zeros(H)
                c = np.zeros(H)
                ts = range(int(seq_len[b]))
                for t in (reversed(ts) if rev else ts):
                    z = np.concatenate([x[t, b], h]) @ p["kernel"] + p["bias"]
                    i, g, f, o = z[:H], z[H:2 * H], z[2 * H:3 * H], z[3 * H:]
                    if "w_i_diag" in p:
                        c_new = sig(f + forget_bias + p["w_f_diag"] * c) * c + sig(i + p["w_i_diag"] * c) * np.tanh(g)
                    else:
                        c_new = sig(f + forget_bias) * c + sig(i) * np.tanh(g)
                    if cell_clip is not None:
                        c_new = np.clip(c_new, -cell_clip, cell_clip)
                    if "w_o_diag" in p:
                        h = sig(o + p["w_o_diag"] * c_new) * np.tanh(c_new)
                    else:
                        h = sig(o) * np.tanh(c_new)
                    c = c_new
                    out[t, b] = h
            outs.append(out)
        x = np.concatenate(outs, axis=2)
    return x


# ---------------------------------------------------------------- GRU (models/encoders/core/gru.py)
def gru_cell_step(x, h_prev, p):
    """tf.contrib.rnn.GRUCell (TF 1.x): [r, u] = sigmoid([x, h] W_g + b_g); c = tanh([x, r*h] W_c + b_c);
    h' = u*h + (1-u)*c.  p: 'gates/kernel' [(D+H),2H], 'gates/bias', 'candidate/kernel' [(D+H),H], 'candidate/bias'."""
    H = p["candidate/bias"].shape[0]
    ru = torch.sigmoid(torch.cat([x, h_prev], 1) @ p["gates/kernel"] + p["gates/bias"])
    r, u = ru[:, :H], ru[:, H:]
    c = torch.tanh(torch.cat([x, r * h_prev], 1) @ p["candidate/kernel"] + p["candidate/bias"])
    return u * h_prev + (1 - u) * c


def _run_gru_direction(x_tbd, seq_len, p, reverse):
    T, B, _ = x_tbd.shape
    H = p["candidate/bias"].shape[0]
    h = x_tbd.new_zeros(B, H)
    outs = [None] * T
    lens = torch.as_tensor(np.asarray(seq_len), dtype=torch.long)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        m = (t < lens).to(x_tbd.dtype).unsqueeze(1)
        h_new = gru_cell_step(x_tbd[t], h, p)
        outs[t] = h_new * m
        h = h_new * m + h * (1 - m)
    return torch.stack(outs, 0), h


def gru_forward(x_btd, seq_len, layers, bidirectional):
    """layers: list of {'fw': params[, 'bw': params]}.  -> (outputs [T,B,H or 2H] time-major, final state of the last
    layer: (h_fw, h_bw) or h)."""
    x = x_btd.transpose(0, 1)
    final = None
    for layer in layers:
        of, hf = _run_gru_direction(x, seq_len, layer["fw"], False)
        if bidirectional:
            ob, hb = _run_gru_direction(x, seq_len, layer["bw"], True)
            x, final = torch.cat([of, ob], 2), (hf, hb)
        else:
            x, final = of, hf
    return x, final
