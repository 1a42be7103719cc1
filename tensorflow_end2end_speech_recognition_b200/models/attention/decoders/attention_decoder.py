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

Forward only (greedy inference and teacher-forced logits).  The decoder backward pass and
dropout inside the decoder are not built yet -- ``keep_prob`` other than 1 raises.
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
    def __call__(self, initial_state, helper, keep_prob=1.0):
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
        self.attention_layer.precompute_keys(enc)
        w_av = v["attentional_vector/weights"]
        p = ops._ptr
        n_steps = 0
        t = 0
        while t < L:
            z = ops.gemm(xh, cv["kernel"])
            c_new, h_new = ops.lstm_cell_pointwise(z, cv["bias"], peep, c_state,
                                                   self.rnn_cell.forget_bias, self.rnn_cell.clip_cell)
            alpha, ctx = self.attention_layer(enc, h_new, self.encoder_outputs_seq_len, prev_alpha)
            av = ops.gemm(h_new, w_av[:Hd])
            ops.gemm(ctx, w_av[Hd:], out=av, beta=1.0)
            ops.tanh_(av)
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
        return outs, LSTMStateTuple(c_state, h_state)

    @staticmethod
    def _greedy_length(ids, eos, n_run):
        """Number of iterations the reference loop runs: one past the step at which the last
        unfinished row emitted <EOS> (or the iteration cap)."""
        hit = (ids == eos)
        any_hit = hit.any(dim=1)
        first = torch.where(any_hit, hit.to(torch.int32).argmax(dim=1) + 1,
                            torch.full_like(any_hit, n_run, dtype=torch.int64))
        return int(min(n_run, int(first.max().item())))
