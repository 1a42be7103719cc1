"""Attention decoder -- host mirror of ``models/attention/decoders/attention_decoder.py``.

Same constructor arguments, ``__call__(initial_state, helper) -> (outputs, final_state)`` and
``AttentionDecoderOutput`` fields as the reference (:18-25, :103-141).  One iteration of the
reference's ``dynamic_decode`` loop (dynamic_decoder.py:148-196) is
    cell pre-activation GEMM -> b2_lstm_cell_pointwise -> attention (query GEMM +
    b2_attention_step_forward) -> attentional-vector GEMMs + b2_tanh_inplace -> logits GEMM
    -> b2_argmax_rows -> b2_decoder_step_emit
issued back to back by ONE native call (``b2_attention_decoder_forward``, csrc/decoder_loop.cu;
the reference's loop also runs inside the framework runtime, not in Python) with no host
synchronisation in the teacher-forced path; in the greedy path the all-finished test of the loop
condition (:143-146) is polled every ``poll_every`` iterations and the outputs are cut at the
first all-finished step, which is what the reference returns because finished rows only emit
zeros.

Decoder-cell output dropout and label-embedding dropout act in the training pass (counter-hash
masks, same convention as the encoder's).  ``backward(dlogits)`` is the teacher-forced loop differentiated: everything that is not
sequential is time-batched (output layer, attentional vector, cell-kernel / embedding /
W_query / W_keys gradients, d(enc) through the context as one GEMM per utterance); the
per-step remainder is attention backward -> query GEMM -> cell gate math backward -> cell
kernel GEMM.
"""
from collections import namedtuple

import ctypes as C_

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
    def _desc(self, B, T, E, emb, keep_prob_decoder=1.0, keep_prob_embedding=1.0, dropout_seed=0):
        al = self.attention_layer
        t = al.attention_type
        Hd = self.rnn_cell.num_units
        loc = t in ("hybrid", "location")
        return _lib.DecoderDesc(
            B, T, E, Hd, self._query_width(Hd, E), emb, self.num_classes,
            1 if t in ("dot_product", "luong_dot", "luong_general") else 0,
            1 if al.query_is_projected() else 0,
            int(al.variables["filter"].shape[0]) if loc else 0,
            float(al.sharpening_factor), int(bool(al.sigmoid_smoothing)),
            float(self.rnn_cell.forget_bias), float(self.rnn_cell.clip_cell or 0.0),
            int(bool(self.feed_previous_attention)), float(keep_prob_decoder), float(keep_prob_embedding),
            int(dropout_seed))

    def _param_struct(self, cell, att, dec, embedding, E):
        """b2_decoder_params / b2_decoder_grads from the variable (or gradient) dicts"""
        s = _lib.DecoderParams()
        ptr = lambda x: x.data_ptr() if x is not None else None
        t = self.attention_layer.attention_type
        s.cell_kernel, s.cell_bias = ptr(cell["kernel"]), ptr(cell["bias"])
        s.w_i_diag, s.w_f_diag, s.w_o_diag = (ptr(cell.get(k)) for k in ("w_i_diag", "w_f_diag", "w_o_diag"))
        if t == "luong_concat":
            s.w_query = ptr(att["W_concat/weights"][E:])
        elif "W_query/weights" in att:
            s.w_query = ptr(att["W_query/weights"])
        s.conv_filter, s.w_filter = ptr(att.get("filter")), ptr(att.get("W_filter/weights"))
        s.b_filter, s.v_a = ptr(att.get("W_filter/biases")), ptr(att.get("v_a"))
        s.w_av, s.w_out = ptr(dec["attentional_vector/weights"]), ptr(dec["output_layer/weights"])
        s.b_out, s.embedding = ptr(dec["output_layer/biases"]), ptr(embedding)
        return s

    def __call__(self, initial_state, helper, keep_prob=1.0, is_training=False, keep_prob_embedding=1.0,
                 dropout_seed=0):
        """keep_prob: DropoutWrapper(output_keep_prob) of the decoder cell (attention_seq2seq.py:367-369);
        keep_prob_embedding: dropout on the embedded labels (:438-439).  Both act in the training pass only."""
        lib = _lib.load()
        enc = self.encoder_outputs.contiguous()
        B, T, E = enc.shape
        dev = enc.device
        Hd, C = self.rnn_cell.num_units, self.num_classes
        emb_table = helper.embedding.contiguous()
        emb = emb_table.shape[1]
        teacher = isinstance(helper, TrainingHelper)
        if teacher:
            labels, dec_len = helper.labels, helper.sequence_length
            L = labels.shape[1] - 1
            sos = eos = -1
        else:
            assert isinstance(helper, GreedyEmbeddingHelper)
            if self.max_decode_length is None:
                raise ValueError("greedy decoding needs max_decode_length")
            if is_training:
                raise ValueError("is_training needs a TrainingHelper")
            labels = dec_len = None
            L = int(self.max_decode_length)
            sos, eos = int(helper.start_tokens[0].item()), helper.end_token
        Ls = max(L, 1)
        f32 = dict(dtype=torch.float32, device=dev)
        out_logits = torch.zeros((B, Ls, C), **f32)
        out_ids = torch.zeros((B, Ls), dtype=torch.int32, device=dev)
        out_av = torch.zeros((B, Ls, Hd), **f32)
        out_alpha = torch.zeros((B, Ls, T), **f32)
        out_ctx = torch.zeros((B, Ls, E), **f32)
        c_state, h_state = torch.empty((B, Hd), **f32), torch.empty((B, Hd), **f32)
        finished = torch.empty(B, dtype=torch.int32, device=dev)
        keys = self.attention_layer.precompute_keys(enc)
        desc = self._desc(B, T, E, emb, keep_prob if is_training else 1.0,
                          keep_prob_embedding if is_training else 1.0, dropout_seed)
        ps = self._param_struct(self.cell_variables, self.attention_layer.variables, self.variables, emb_table, E)
        reserve = None
        if is_training:
            reserve = torch.empty(lib.b2_attention_decoder_reserve_bytes(C_.byref(desc), L), dtype=torch.uint8,
                                  device=dev)
        nbytes = lib.b2_attention_decoder_workspace_bytes(C_.byref(desc), L)
        ws = ops.workspace("decoder", nbytes, dev)
        steps = C_.c_int32(0)
        p = ops._ptr
        rc = lib.b2_attention_decoder_forward(
            C_.byref(desc), C_.byref(ps), p(enc), p(keys), p(self.encoder_outputs_seq_len),
            p(initial_state.c.contiguous()), p(initial_state.h.contiguous()), p(labels),
            labels.shape[1] if teacher else 0, p(dec_len), sos, eos, L, 0 if teacher else self.poll_every,
            p(reserve), p(out_logits), p(out_ids), p(out_av), p(out_alpha), p(out_ctx), p(c_state), p(h_state),
            p(finished), C_.byref(steps), p(ws), nbytes, ops._stream())
        _lib.check(rc, "b2_attention_decoder_forward")
        n_steps = int(steps.value)
        if L <= 0:
            c_state, h_state = initial_state.c.clone(), initial_state.h.clone()
        # cut at the first all-finished step: later rows are zeros and carry no information
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
        self._saved = None
        if is_training:
            self._saved = {"reserve": reserve, "desc": desc, "L": L, "labels": labels, "enc": enc,
                           "emb_table": emb_table, "keys": keys}
        return outs, LSTMStateTuple(c_state, h_state)

    def beam_search(self, initial_state, embedding, sos_index, eos_index, beam_width, length_penalty_weight=0.6):
        """Beam search for every utterance of the batch (reference semantics:
        beam_search/beam_search_decoder.py:234-332, one beam per utterance instead of batch size 1).
        -> (predicted_ids [B,W,L'] int32, lengths [B,W], log_probs [B,W], scores [B,W]); beam 0 is best."""
        lib = _lib.load()
        enc = self.encoder_outputs.contiguous()
        B, T, E = enc.shape
        dev = enc.device
        W, L = int(beam_width), int(self.max_decode_length)
        emb_table = embedding.contiguous()
        keys = self.attention_layer.precompute_keys(enc)
        desc = self._desc(B, T, E, emb_table.shape[1])
        ps = self._param_struct(self.cell_variables, self.attention_layer.variables, self.variables, emb_table, E)
        out_ids = torch.empty((B, W, L), dtype=torch.int32, device=dev)
        out_len = torch.empty((B, W), dtype=torch.int32, device=dev)
        out_lp = torch.empty((B, W), dtype=torch.float32, device=dev)
        out_sc = torch.empty((B, W), dtype=torch.float32, device=dev)
        nbytes = lib.b2_attention_decoder_beam_workspace_bytes(C_.byref(desc), W, L)
        ws = ops.workspace("decoder_beam", nbytes, dev)
        steps = C_.c_int32(0)
        p = ops._ptr
        rc = lib.b2_attention_decoder_beam_search(
            C_.byref(desc), C_.byref(ps), p(enc), p(keys), p(self.encoder_outputs_seq_len),
            p(initial_state.c.contiguous()), p(initial_state.h.contiguous()), int(sos_index), int(eos_index), W,
            float(length_penalty_weight), L, self.poll_every, p(out_ids), p(out_len), p(out_lp), p(out_sc),
            C_.byref(steps), p(ws), nbytes, ops._stream())
        _lib.check(rc, "b2_attention_decoder_beam_search")
        return out_ids[:, :, :int(steps.value)], out_len, out_lp, out_sc

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
        lib = _lib.load()
        al = self.attention_layer
        enc, desc, L = sv["enc"], sv["desc"], sv["L"]
        B, T, E = enc.shape
        Hd = self.rnn_cell.num_units
        dev = enc.device
        f32 = dict(dtype=torch.float32, device=dev)
        t_ = al.attention_type
        if t_ == "luong_dot":
            d_keys = d_enc                                                    # keys are the encoder states
        elif t_ == "location":
            d_keys = None
        else:
            d_keys = torch.zeros((B, T, sv["keys"].shape[-1]), **f32)
        dl_tm = ops.transpose_01(dlogits.contiguous())                        # [L,B,C]
        ps = self._param_struct(self.cell_variables, al.variables, self.variables, sv["emb_table"], E)
        gs = self._param_struct(cell_grads, att_grads, grads, emb_grad, E)
        dc0, dh0 = torch.empty((B, Hd), **f32), torch.empty((B, Hd), **f32)
        nbytes = lib.b2_attention_decoder_workspace_bytes(C_.byref(desc), L)
        ws = ops.workspace("decoder", nbytes, dev)
        p = ops._ptr
        rc = lib.b2_attention_decoder_backward(
            C_.byref(desc), C_.byref(ps), p(enc), p(sv["keys"]), p(self.encoder_outputs_seq_len), p(sv["labels"]),
            sv["labels"].shape[1], L, p(sv["reserve"]), p(dl_tm), C_.byref(gs), p(d_keys), p(d_enc), p(dc0), p(dh0),
            p(ws), nbytes, ops._stream())
        _lib.check(rc, "b2_attention_decoder_backward")
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
