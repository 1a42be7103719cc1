"""Attention seq2seq / joint CTC-attention training loss, CPU restatement in torch (float64,
autograd gives the gradients the GPU backward kernels are checked against).
TEST INFRASTRUCTURE.

Follows ``models/attention/attention_seq2seq.py`` ``_build`` :196-283 (encoder -> bridge ->
teacher-forced decoder), ``_decode_train`` :413-459, ``compute_loss`` :579-664
(``logits / logits_temperature`` :270, ``tf.contrib.seq2seq.sequence_loss`` with
``sequence_mask(labels_seq_len - 1)`` weights on ``labels[:, 1:]`` :619-636, weight decay over
every variable without 'bias' in its name :607-613) and
``models/attention/joint_ctc_attention.py`` ``ctc_logits`` :182-235, ``compute_loss`` :268-322
(``lambda * mean(ctc) + (1 - lambda) * sequence_loss``).

TF-upstream fact restated: ``sequence_loss(average_across_timesteps=True,
average_across_batch=True)`` = sum(xent * w) / (sum(w) + 1e-12) over all (b, t).

The decoder loop is the torch twin of ``oracle/attention_decoder.py::decode`` (teacher
forcing only) and ``attention_step_t`` the torch twin of ``oracle/attention.py``; the CPU test
suite checks the twins against each other.
"""
import numpy as np
import torch

from . import lstm as olstm
from .attention import FLOAT32_MIN


def _conv1d_same_t(x_bt, filt_k10):
    B, T = x_bt.shape
    k = filt_k10.shape[0]
    pl = (k - 1) // 2
    xp = torch.zeros((B, T + k - 1), dtype=x_bt.dtype)
    xp[:, pl:pl + T] = x_bt
    win = xp.unfold(1, k, 1)                                  # [B, T, k]
    return win @ filt_k10                                     # [B, T, 10]


def attention_step_t(attention_type, enc, query, enc_len, prev_alpha, p, sharpening_factor=1.0,
                     sigmoid_smoothing=False):
    """torch twin of oracle.attention.attention_step -> (weights [B,T], context [B,E])."""
    B, T, E = enc.shape
    t = attention_type
    if t in ("bahdanau_content", "location", "hybrid", "dot_product"):
        wq = query @ p["W_query/weights"]
        wk = enc @ p["W_keys/weights"]
        if t != "dot_product":
            wk = wk + p["W_keys/biases"]
        if t == "dot_product":
            e = torch.einsum("bta,ba->bt", wk, wq)
        elif t == "bahdanau_content":
            e = (p["v_a"] * torch.tanh(wk + wq[:, None, :])).sum(2)
        else:
            f = _conv1d_same_t(prev_alpha, p["filter"][:, 0, :])
            wf = f @ p["W_filter/weights"] + p["W_filter/biases"]
            if t == "hybrid":
                e = (p["v_a"] * torch.tanh(wk + wq[:, None, :] + wf)).sum(2)
            else:
                e = (p["v_a"] * torch.tanh(wq[:, None, :] + wf)).sum(2)
    elif t == "luong_dot":
        e = torch.einsum("bte,be->bt", enc, query)
    elif t == "luong_general":
        e = torch.einsum("btd,bd->bt", enc @ p["W_keys/weights"], query)
    elif t == "luong_concat":
        cat = torch.cat([enc, query[:, None, :].expand(B, T, query.shape[1])], dim=2)
        e = (p["v_a"] * torch.tanh(cat @ p["W_concat/weights"])).sum(2)
    else:
        raise ValueError(attention_type)
    mask = (torch.arange(T)[None, :] < torch.as_tensor(np.asarray(enc_len))[:, None]).to(enc.dtype)
    e = e * mask + (1.0 - mask) * FLOAT32_MIN
    e = e * sharpening_factor
    if sigmoid_smoothing:
        s = torch.sigmoid(e) * mask
        w = s / s.sum(-1, keepdim=True)
    else:
        w = torch.softmax(e, dim=-1)
    return w, torch.einsum("bt,bte->be", w, enc)


def bridge_t(final_state, p):
    flat = torch.cat([s for pair in final_state for s in pair], dim=1)
    out = flat @ p["bridge/weights"] + p["bridge/biases"]
    Hd = out.shape[1] // 2
    return out[:, :Hd], out[:, Hd:]


def decode_train_t(p, attention_type, enc, enc_len, initial_state, labels, labels_seq_len,
                   sharpening_factor=1.0, sigmoid_smoothing=False, feed_previous_attention=False,
                   cell_clip=None, emb_mask=None, keep_prob_embedding=1.0, dec_mask=None, keep_prob_decoder=1.0):
    """Teacher-forced decoder -> dict(logits [B,L,C], predicted_ids, attention_weights).
    emb_mask [B,T_out,emb] / dec_mask [L,B,Hd]: {0,1} dropout masks of the embedded labels
    (attention_seq2seq.py:438-439) and of the decoder cell OUTPUT (DropoutWrapper, :367-369: the state that
    recurs is not dropped)."""
    B, T, E = enc.shape
    labels = np.asarray(labels)
    emb = p["W_embedding"]
    embedded = emb[torch.as_tensor(labels, dtype=torch.long)]               # [B, T_out, emb]
    if emb_mask is not None and keep_prob_embedding < 1.0:
        embedded = embedded * torch.as_tensor(emb_mask, dtype=embedded.dtype) / keep_prob_embedding
    dec_in = embedded[:, :-1]                                               # [B, L, emb]
    L = dec_in.shape[1]
    seq = np.asarray(labels_seq_len) - 1
    c, h = initial_state
    ctx = enc.new_zeros((B, E))
    alpha_state = enc.new_zeros((B, T))
    finished = seq <= 0
    logits_all, ids_all, alpha_all = [], [], []
    time = 0
    while not finished.all():
        x = torch.cat([dec_in[:, time], ctx], dim=1)
        h_new, c_new = olstm.lstm_cell_step(x, h, c, p["cell"], 1.0, cell_clip)
        h_out = h_new
        if dec_mask is not None and keep_prob_decoder < 1.0:
            h_out = h_new * torch.as_tensor(dec_mask[time], dtype=h_new.dtype) / keep_prob_decoder
        alpha, ctx = attention_step_t(attention_type, enc, h_out, enc_len,
                                      alpha_state if feed_previous_attention else enc.new_zeros((B, T)),
                                      p["attention"], sharpening_factor, sigmoid_smoothing)
        av = torch.tanh(torch.cat([h_out, ctx], dim=1) @ p["attentional_vector/weights"])
        logits = av @ p["output_layer/weights"] + p["output_layer/biases"]
        keep = torch.as_tensor(~finished)
        m = keep[:, None].to(enc.dtype)
        logits_all.append(logits * m)
        ids_all.append(torch.where(keep, logits.argmax(1), torch.zeros(B, dtype=torch.long)))
        alpha_all.append(alpha * m)
        c = torch.where(keep[:, None], c_new, c)
        h = torch.where(keep[:, None], h_new, h)
        alpha_state = alpha
        finished = finished | ((time + 1) >= seq)
        time += 1
    assert time == L or time == int(seq.max()), (time, L)
    return {"logits": torch.stack(logits_all, 1), "predicted_ids": torch.stack(ids_all, 1),
            "attention_weights": torch.stack(alpha_all, 1), "final_state": (c, h)}


def sequence_loss_t(logits, targets, lengths):
    """logits [B,L,C], targets [B,L] int, lengths [B] -> scalar (see module docstring)."""
    B, L, C = logits.shape
    w = (torch.arange(L)[None, :] < torch.as_tensor(np.asarray(lengths))[:, None]).to(logits.dtype)
    logp = torch.log_softmax(logits, dim=-1)
    tgt = torch.as_tensor(np.asarray(targets), dtype=torch.long)
    xent = -logp.gather(2, tgt[:, :, None])[:, :, 0]
    return (xent * w).sum() / (w.sum() + 1e-12)


def split_variables(variables, num_layers, use_peephole_enc=True):
    """flat TF-name dict (torch tensors) -> the nested dicts the oracle functions take."""
    from .model import layers_from_variables
    enc_vars = {k[len("encoder/"):]: v for k, v in variables.items() if k.startswith("encoder/")}
    layers = layers_from_variables(enc_vars, num_layers, use_peephole_enc)
    pre_c = "decoder/decoder_rnn_cell/lstm_cell/"
    pre_a = "decoder/attention_decoder/attention_layer/"
    p = {"cell": {k[len(pre_c):]: v for k, v in variables.items() if k.startswith(pre_c)},
         "attention": {k[len(pre_a):]: v for k, v in variables.items() if k.startswith(pre_a)},
         "W_embedding": variables["decoder/output_embedding/W_embedding"],
         "bridge/weights": variables["decoder/bridge/weights"],
         "bridge/biases": variables["decoder/bridge/biases"]}
    for k in ("attentional_vector/weights", "output_layer/weights", "output_layer/biases"):
        p[k] = variables["decoder/attention_decoder/" + k]
    return layers, p


def seq2seq_loss(variables, cfg, inputs_btd, inputs_seq_len, labels, labels_seq_len,
                 ctc_labels=None):
    """variables: flat dict of torch tensors (requires_grad as wanted).  cfg keys:
    num_layers, attention_type, use_peephole, sharpening_factor, sigmoid_smoothing,
    logits_temperature, weight_decay, feed_previous_attention, lambda_weight (joint only),
    ctc_faithful_reshape (joint: reproduce the [B*T] -> [T,B] reshape of the reference).
    Returns dict(total_loss, sequence_loss, ctc_loss, logits, ctc_logits, encoder_outputs)."""
    layers, p = split_variables(variables, cfg["num_layers"], cfg.get("use_peephole", True))
    enc_tm, final = olstm.blstm_forward(inputs_btd, inputs_seq_len, layers)
    enc = enc_tm.transpose(0, 1)                                             # batch-major
    init = bridge_t(final, p)
    dec = decode_train_t(p, cfg["attention_type"], enc, inputs_seq_len, init, labels, labels_seq_len,
                         cfg.get("sharpening_factor", 1.0), cfg.get("sigmoid_smoothing", False),
                         cfg.get("feed_previous_attention", False), None, cfg.get("emb_mask"),
                         cfg.get("keep_prob_embedding", 1.0), cfg.get("dec_mask"), cfg.get("keep_prob_decoder", 1.0))
    logits = dec["logits"] / cfg.get("logits_temperature", 1.0) + 1e-10
    labels = np.asarray(labels)
    L = logits.shape[1]
    seq = sequence_loss_t(logits, labels[:, 1:1 + L], np.asarray(labels_seq_len) - 1)
    out = {"sequence_loss": seq, "logits": logits, "encoder_outputs": enc, "decoder": dec}
    total = seq
    lam = cfg.get("lambda_weight")
    if lam is not None:
        B, T, E2 = enc.shape
        w, b = variables["ctc_output/weights"], variables["ctc_output/biases"]
        if cfg.get("ctc_faithful_reshape", False):
            ctc_logits = (enc.reshape(B * T, E2) @ w + b).reshape(T, B, -1)  # joint_ctc_attention.py:223-226
        else:
            ctc_logits = enc_tm @ w + b
        C = ctc_logits.shape[-1]
        lens = torch.tensor([len(l) for l in ctc_labels], dtype=torch.long)
        flat = torch.tensor([v for l in ctc_labels for v in l], dtype=torch.long)
        ilen = torch.as_tensor(np.asarray(inputs_seq_len), dtype=torch.long)
        if bool((lens > ilen).any()):
            raise ValueError("Not enough time for target transition sequence")   # ignore_longer=False
        ctc = torch.nn.functional.ctc_loss(torch.log_softmax(ctc_logits, -1), flat, ilen, lens, blank=C - 1,
                                           reduction="none", zero_infinity=False).mean()
        out["ctc_loss"], out["ctc_logits"] = ctc, ctc_logits
        total = lam * ctc + (1.0 - lam) * seq
    wd = cfg.get("weight_decay", 0.0)
    if wd > 0:
        total = total + wd * sum(0.5 * (v ** 2).sum() for k, v in variables.items() if "bias" not in k.lower())
    out["total_loss"] = total
    return out
