"""Attention decoder -- host mirror of ``models/attention/decoders/attention_decoder.py``.

Same constructor arguments, ``__call__(initial_state, helper) -> (outputs, final_state)`` and
``AttentionDecoderOutput`` fields as the reference (:18-25, :103-141).  One iteration of the
reference's ``dynamic_decode`` loop (dynamic_decoder.py:148-196) is
    cell pre-activation GEMM -> b2_lstm_cell_pointwise -> attention (query GEMM +
    b2_attention_step_forward) -> attentional-vector GEMMs + b2_tanh_inplace -> logits GEMM
    -> b2_argmax_rows -> b2_decoder_step_emit
with no host synchronisation; the all-finished test of the loop condition (:143-146) is polled
every ``poll_every`` iterations and the outputs are cut at the first all-finished step, which
is what the reference returns because finished rows only emit zeros.

``backward(dlogits)`` is the teacher-forced loop differentiated: everything that is not
sequential is time-batched (output layer, attentional vector, cell-kernel / embedding /
W_query / W_keys gradients, d(enc) through the context as one GEMM per utterance); the
per-step remainder is attention backward -> query GEMM -> cell gate math backward -> cell
kernel GEMM.  Dropout inside the decoder is not built -- ``keep_prob`` other than 1 raises.
"""
from collections import namedtuple

import numpy as np
import torch

from .... import _lib, ops
from ..bridge import LSTMStateTuple
from .helpers import GreedyEmbeddingHelper, TrainingHelper

AttentionDecoderOutput = namedtuple(
    "AttentionDecoderOutput",
    ["logits", "predicted_ids", "decoder_output", "attention_weights", "context_vector"])


class LSTMBlockCell(object):
    """Configuration of the decoder cell (attention_seq2seq.py:352-363)."""

    def __init__(self, num_units, forget_bias=1.0, clip_cell=None, use_peephole=False):
        self.num_units, self.forget_bias = num_units, forget_bias
        self.clip_cell, self.use_peephole = clip_cell, use_peephole

    @property
    def state_size(self):
        return LSTMStateTuple(self.num_units, self.num_units)

    @property
    def output_size(self):
        return self.num_units

    def create_variables(self, input_size, parameter_init, rng, device):
        a, H = parameter_init, self.num_units
        v = {"kernel": rng.uniform(-a, a, (input_size + H, 4 * H)).astype(np.float32),
             "bias": np.zeros(4 * H, np.float32)}
        if self.use_peephole:
            for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
                v[k] = rng.uniform(-a, a, H).astype(np.float32)
        return {k: torch.tensor(x, device=device) for k, x in v.items()}


class AttentionDecoder(object):
    def __init__(self, rnn_cell, parameter_init, max_decode_length, num_classes, encoder_outputs,
                 encoder_outputs_seq_len, attention_layer, time_major, mode=None,
                 name="attention_decoder", feed_previous_attention=False, poll_every=8):
        self.rnn_cell = rnn_cell
        self.parameter_init = parameter_init
        self.max_decode_length = max_decode_length
        self.num_classes = num_classes
        self.encoder_outputs = encoder_outputs                      # [B, T, E] batch-major
        self.encoder_outputs_seq_len = encoder_outputs_seq_len
        self.attention_layer = attention_layer
        self.time_major = time_major
        self.mode = mode
        self.name = name
        # the reference's loop body always sees the zero weights of initialize()
        self.feed_previous_attention = feed_previous_attention
        self.poll_every = poll_every
        self.variables = None
        self.cell_variables = None

    # ------------------------------------------------------------ variables
    def create_variables(self, embedding_dim, rng, device):
        E = self.encoder_outputs.shape[-1] if self.encoder_outputs is not None else None
        return self._create_variables(embedding_dim, E, rng, device)

    def _create_variables(self, embedding_dim, E, rng, device):
        Hd, std = self.rnn_cell.num_units, self.parameter_init

        def tn(shape):
            x = rng.normal(0, std, size=shape)
            bad = np.abs(x) > 2 * std
            while bad.any():
                x[bad] = rng.normal(0, std, size=int(bad.sum()))
                bad = np.abs(x) > 2 * std
            return torch.tensor(x.astype(np.float32), device=device)
        self.cell_variables = self.rnn_cell.create_variables(embedding_dim + E, std, rng, device)
        self.variables = {"attentional_vector/weights": tn((Hd + E, Hd)),
                          "output_layer/weights": tn((Hd, self.num_classes)),
                          "output_layer/biases": torch.zeros(self.num_classes, device=device)}
        return self.variables

    # ----------------------------------------------------------------- loop
    def __call__(self, initial_state, helper, keep_prob=1.0, is_training=False):
        if keep_prob != 1.0:
            raise NotImplementedError("decoder dropout: forward-only decoder, keep_prob must be 1")
        lib = _lib.load()
        enc = self.encoder_outputs.contiguous()
        B, T, E = enc.shape
        dev = enc.device
        Hd, C = self.rnn_cell.num_units, self.num_classes
        emb_table = helper.embedding.contiguous()
        emb = emb_table.shape[1]
        cv, v = self.cell_variables, self.variables
        peep = (cv["w_i_diag"], cv["w_f_diag"], cv["w_o_diag"]) if "w_i_diag" in cv else None
        teacher = isinstance(helper, TrainingHelper)
        if teacher:
            labels, dec_len = helper.labels, helper.sequence_length
            L = labels.shape[1] - 1
            max_iter = 0
            first_ids = labels[:, 0].contiguous()
            finished = (dec_len <= 0).to(torch.int32)
        else:
            assert isinstance(helper, GreedyEmbeddingHelper)
            if self.max_decode_length is None:
                raise ValueError("greedy decoding needs max_decode_length")
            labels = dec_len = None
            L = max_iter = int(self.max_decode_length)
            first_ids = helper.start_tokens.to(torch.int32)
            finished = torch.zeros(B, dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        out_logits = torch.zeros((B, max(L, 1), C), **f32)
        out_ids = torch.zeros((B, max(L, 1)), dtype=torch.int32, device=dev)
        out_av = torch.zeros((B, max(L, 1), Hd), **f32)
        out_alpha = torch.zeros((B, max(L, 1), T), **f32)
        out_ctx = torch.zeros((B, max(L, 1), E), **f32)
        c_state = initial_state.c.clone().contiguous()
        h_state = initial_state.h.clone().contiguous()
        # first cell input: [emb(first id) ; zero context ; h0]      (attention_decoder.py:143-168)
        xh = torch.zeros((B, emb + E + Hd), **f32)
        xh[:, :emb] = emb_table[first_ids.long()]
        xh[:, emb + E:] = h_state
        zeros_alpha = torch.zeros((B, T), **f32)
        prev_alpha = zeros_alpha
        keys = self.attention_layer.precompute_keys(enc)
        w_av = v["attentional_vector/weights"]
        p = ops._ptr
        save = None
        if is_training:
            if not teacher:
                raise ValueError("is_training needs a TrainingHelper")
            if self.feed_previous_attention:
                raise NotImplementedError("training with feed_previous_attention=True is not built "
                                          "(the reference never feeds previous weights, SURVEY A.7.1)")
            A = self._query_width(Hd, E)
            Ls = max(L, 1)
            save = {"xh": torch.zeros((Ls + 1, B, emb + E + Hd), **f32), "z": torch.empty((Ls, B, 4 * Hd), **f32),
                    "c": torch.empty((Ls + 1, B, Hd), **f32), "h": torch.empty((Ls, B, Hd), **f32),
                    "alpha": torch.empty((Ls, B, T), **f32), "ctx": torch.empty((Ls, B, E), **f32),
                    "av": torch.empty((Ls, B, Hd), **f32),
                    "q": torch.empty((Ls, B, A), **f32) if self.attention_layer.query_is_projected() else None,
                    "energy": torch.empty((Ls, B, T), **f32) if self.attention_layer.sigmoid_smoothing else None,
                    "L": L, "labels": labels, "dec_len": dec_len, "emb_table": emb_table, "enc": enc}
            save["xh"][0].copy_(xh)
            save["c"][0].copy_(c_state)
            xh = save["xh"][0]
        n_steps = 0
        t = 0
        while t < L:
            if save is not None:
                z = ops.gemm(xh, cv["kernel"], out=save["z"][t])
                c_new, h_new = ops.lstm_cell_pointwise(z, cv["bias"], peep, c_state, self.rnn_cell.forget_bias,
                                                       self.rnn_cell.clip_cell, out_c=save["c"][t + 1],
                                                       out_h=save["h"][t])
                att_out = {"alpha": save["alpha"][t], "context": save["ctx"][t],
                           "q": save["q"][t] if save["q"] is not None else None,
                           "energy": save["energy"][t] if save["energy"] is not None else None}
                alpha, ctx = self.attention_layer(enc, h_new, self.encoder_outputs_seq_len, prev_alpha, out=att_out)
                av = ops.gemm(h_new, w_av[:Hd], out=save["av"][t])
            else:
                z = ops.gemm(xh, cv["kernel"])
                c_new, h_new = ops.lstm_cell_pointwise(z, cv["bias"], peep, c_state,
                                                       self.rnn_cell.forget_bias, self.rnn_cell.clip_cell)
                alpha, ctx = self.attention_layer(enc, h_new, self.encoder_outputs_seq_len, prev_alpha)
                av = ops.gemm(h_new, w_av[:Hd])
            ops.gemm(ctx, w_av[Hd:], out=av, beta=1.0)
            ops.tanh_(av)
            if save is not None:
                xh = save["xh"][t + 1]
            logits = ops.gemm(av, v["output_layer/weights"], bias=v["output_layer/biases"])
            ids = ops.argmax_rows(logits)
            rc = lib.b2_decoder_step_emit(
                B, C, Hd, E, T, emb, t, max(L, 1), p(logits), p(ids), p(av), p(alpha), p(ctx),
                p(c_new), p(h_new), p(c_state), p(h_state), p(finished), p(emb_table),
                p(labels), labels.shape[1] if teacher else 0, p(dec_len),
                -1 if teacher else helper.end_token, max_iter, p(xh), p(out_logits), p(out_ids),
                p(out_av), p(out_alpha), p(out_ctx), ops._stream())
            _lib.check(rc, "b2_decoder_step_emit")
            if self.feed_previous_attention:
                prev_alpha = alpha
            t += 1
            n_steps = t
            if t % self.poll_every == 0 and bool(finished.all().item()):
                break
        # cut at the first all-finished step: rows of zeros in out_ids/out_logits past it carry
        # no information, find it from the per-step "any row still emitting" flags
        if n_steps > 0:
            if teacher:
                n_steps = int(min(L, max(int(dec_len.max().item()), 0)))
            else:
                n_steps = self._greedy_length(out_ids[:, :n_steps], helper.end_token, n_steps)
        outs = AttentionDecoderOutput(
            logits=out_logits[:, :n_steps], predicted_ids=out_ids[:, :n_steps],
            decoder_output=out_av[:, :n_steps], attention_weights=out_alpha[:, :n_steps],
            context_vector=out_ctx[:, :n_steps])
        if self.time_major:
            outs = AttentionDecoderOutput(*[x.transpose(0, 1).contiguous() for x in outs])
        self._saved = save
        return outs, LSTMStateTuple(c_state, h_state)

    def _query_width(self, Hd, E):
        al = self.attention_layer
        if al.attention_type in ("bahdanau_content", "location", "hybrid", "dot_product", "luong_concat"):
            return al.num_units
        return Hd                                    # luong_dot / luong_general: the cell output itself

    # -------------------------------------------------------------- backward
    def backward(self, dlogits, grads, cell_grads, att_grads, emb_grad, d_enc):
        """dlogits [B,L,C] (batch-major, zero on masked steps) -> (dc0, dh0) [B,Hd].
        Accumulates into grads (attentional_vector/output_layer), cell_grads, att_grads (same keys as
        the variable dicts), emb_grad [V,emb] and d_enc [B,T,E]."""
        sv = self._saved
        assert sv is not None, "backward needs a forward pass with is_training=True"
        al = self.attention_layer
        enc = sv["enc"]
        B, T, E = enc.shape
        Hd, C, L = self.rnn_cell.num_units, self.num_classes, sv["L"]
        cv, v = self.cell_variables, self.variables
        emb = sv["emb_table"].shape[1]
        dev = enc.device
        f32 = dict(dtype=torch.float32, device=dev)
        peep = (cv["w_i_diag"], cv["w_f_diag"], cv["w_o_diag"]) if "w_i_diag" in cv else None
        # ---- time-batched head: logits FC, tanh, attentional-vector FC (time-major rows)
        dl = ops.transpose_01(dlogits.contiguous()).view(L * B, C)
        av2 = sv["av"][:L].view(L * B, Hd)
        h2 = sv["h"][:L].view(L * B, Hd)
        ctx2 = sv["ctx"][:L].view(L * B, E)
        ops.gemm(av2, dl, True, False, out=grads["output_layer/weights"], beta=1.0)
        ops.colsum(dl, out=grads["output_layer/biases"], accumulate=True)
        d_av = ops.gemm(dl, v["output_layer/weights"], False, True)
        d_pre = ops.tanh_backward(d_av, av2, out=d_av)
        w_av, g_av = v["attentional_vector/weights"], grads["attentional_vector/weights"]
        ops.gemm(h2, d_pre, True, False, out=g_av[:Hd], beta=1.0)
        ops.gemm(ctx2, d_pre, True, False, out=g_av[Hd:], beta=1.0)
        dh_av = ops.gemm(d_pre, w_av[:Hd], False, True).view(L, B, Hd)
        dctx_all = ops.gemm(d_pre, w_av[Hd:], False, True).view(L, B, E)       # becomes total d(ctx_t)
        # ---- sequential part
        A = self._query_width(Hd, E)
        projected = al.query_is_projected()
        t_ = al.attention_type
        wq = None
        if projected:
            wq = al.variables["W_concat/weights"][E:] if t_ == "luong_concat" else al.variables["W_query/weights"]
        dq_all = torch.empty((L, B, A), **f32) if projected else None
        if t_ == "luong_dot":
            d_keys = d_enc                                                    # keys are the encoder states
        elif t_ == "location":
            d_keys = None
        else:
            d_keys = torch.zeros((B, T, al._keys.shape[-1]), **f32)
        dz_all = torch.empty((L, B, 4 * Hd), **f32)
        demb_all = torch.empty((L, B, emb), **f32)
        dh0 = torch.empty((B, Hd), **f32)
        dc_next = None
        kernel = cv["kernel"]
        k_emb, k_ctx, k_h = kernel[:emb], kernel[emb:emb + E], kernel[emb + E:]
        for t in range(L - 1, -1, -1):
            dh_t = dh_av[t]                     # already holds the input-feeding part of cell t+1
            energy_t = sv["energy"][t] if sv["energy"] is not None else None
            if projected:
                al.backward_step(enc, sv["q"][t], sv["alpha"][t], energy_t, self.encoder_outputs_seq_len,
                                 dctx_all[t], d_keys, dq_all[t], att_grads)
                ops.gemm(dq_all[t], wq, False, True, out=dh_t, beta=1.0)
            else:                               # the query is the cell output itself: dq adds into dh
                al.backward_step(enc, sv["h"][t], sv["alpha"][t], energy_t, self.encoder_outputs_seq_len,
                                 dctx_all[t], d_keys, dh_t, att_grads, dq_accumulate=True)
            dz, dc_next = ops.lstm_cell_pointwise_backward(sv["z"][t], cv["bias"], peep, sv["c"][t], dh_t, dc_next,
                                                           self.rnn_cell.forget_bias, self.rnn_cell.clip_cell,
                                                           out_dz=dz_all[t])
            ops.gemm(dz, k_emb, False, True, out=demb_all[t])
            if t > 0:                           # input feeding: [emb ; ctx_{t-1} ; h_{t-1}] was the cell input
                ops.gemm(dz, k_ctx, False, True, out=dctx_all[t - 1], beta=1.0)
                ops.gemm(dz, k_h, False, True, out=dh_av[t - 1], beta=1.0)
            else:
                ops.gemm(dz, k_h, False, True, out=dh0)
        dc0 = dc_next
        ids_tm = sv["labels"][:, :L].t().contiguous().view(-1)              # time-major token ids
        ops.embedding_grad(demb_all, emb, ids_tm, L * B, emb, emb_grad)
        # ---- time-batched tails
        dz2 = dz_all.view(L * B, 4 * Hd)
        ops.gemm(sv["xh"][:L].view(L * B, -1), dz2, True, False, out=cell_grads["kernel"], beta=1.0)
        ops.colsum(dz2, out=cell_grads["bias"], accumulate=True)
        if peep is not None:
            ops.decoder_peephole_grad(dz_all, sv["c"], L, B, Hd, cell_grads["w_i_diag"], cell_grads["w_f_diag"],
                                      cell_grads["w_o_diag"])
        if projected:
            gq = att_grads["W_concat/weights"][E:] if t_ == "luong_concat" else att_grads["W_query/weights"]
            ops.gemm(h2, dq_all.view(L * B, A), True, False, out=gq, beta=1.0)
        # d(enc) through the context: d_enc[b] += Alpha[b]^T . Dctx[b]
        for b in range(B):
            ops.gemm(sv["alpha"][:L, b], dctx_all[:, b], True, False, out=d_enc[b], beta=1.0)
        al.backward_keys(enc, d_keys, d_enc, att_grads)
        self._saved = None
        return dc0, dh0

    @staticmethod
    def _greedy_length(ids, eos, n_run):
        """Number of iterations the reference loop runs: one past the step at which the last
        unfinished row emitted <EOS> (or the iteration cap)."""
        hit = (ids == eos)
        any_hit = hit.any(dim=1)
        first = torch.where(any_hit, hit.to(torch.int32).argmax(dim=1) + 1,
                            torch.full_like(any_hit, n_run, dtype=torch.int64))
        return int(min(n_run, int(first.max().item())))
