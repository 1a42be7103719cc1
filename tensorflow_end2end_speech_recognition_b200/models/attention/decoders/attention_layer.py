"""Attention layer -- host mirror of ``models/attention/decoders/attention_layer.py``.

Same class name, constructor arguments and call signature as the reference
(``AttentionLayer.__call__(encoder_outputs, decoder_output, encoder_outputs_length,
attention_weights) -> (attention_weights [B,T], context_vector [B,E])``, :45-113).  The key
projection ``W_keys . h_enc`` that the reference recomputes inside every decoder step
(:151-159) is computed once per batch by ``precompute_keys`` (tcgen05 / fp32 GEMM); the step
itself is one CUDA launch (``b2_attention_step_forward``).  Forward only in this round: the
decoder loop, its backward pass and the seq2seq loss are the next rows of SURVEY 8.
"""
import ctypes as C

import numpy as np
import torch

from .... import _lib, ops

ATTENTION_TYPE = ["bahdanau_content", "normed_bahdanau_content", "location", "hybrid",
                  "dot_product", "luong_dot", "scaled_luong_dot", "luong_general",
                  "luong_concat", "baidu_attetion"]
_NOT_IMPLEMENTED = ("normed_bahdanau_content", "scaled_luong_dot", "baidu_attetion")   # reference :188,:287,:267


class AttentionLayer(object):
    def __init__(self, attention_type, num_units, parameter_init, sharpening_factor,
                 sigmoid_smoothing, mode=None, name="attention_layer", precision="fp32"):
        if attention_type not in ATTENTION_TYPE:
            raise ValueError("attention type should be one of [%s], you provided %s." %
                             (", ".join(ATTENTION_TYPE), attention_type))
        if attention_type in _NOT_IMPLEMENTED:
            raise NotImplementedError
        self.attention_type = attention_type
        self.num_units = num_units
        self.parameter_init = parameter_init
        self.sharpening_factor = sharpening_factor
        self.sigmoid_smoothing = sigmoid_smoothing
        self.name = name
        self.precision = precision
        self.variables = None
        self._keys = None

    # ------------------------------------------------------------ variables
    def create_variables(self, encoder_num_units, decoder_num_units, rng, device):
        """TF variable names of the reference scopes; truncated_normal(stddev=parameter_init)
        projections (:142-159), v_a glorot-uniform, conv filter as in :197-202 / :237-241."""
        t, A, E, Dq = self.attention_type, self.num_units, encoder_num_units, decoder_num_units

        def tn(shape, std):
            x = rng.normal(0, std, size=shape)
            bad = np.abs(x) > 2 * std
            while bad.any():
                x[bad] = rng.normal(0, std, size=int(bad.sum()))
                bad = np.abs(x) > 2 * std
            return x.astype(np.float32)
        v = {}
        if t in ("bahdanau_content", "location", "hybrid", "dot_product"):
            v["W_query/weights"] = tn((Dq, A), self.parameter_init)
            v["W_keys/weights"] = tn((E, A), self.parameter_init)
            if t != "dot_product":
                v["W_keys/biases"] = np.zeros(A, np.float32)
        if t == "luong_general":
            v["W_keys/weights"] = tn((E, Dq), self.parameter_init)
        if t == "luong_concat":
            v["W_concat/weights"] = tn((E + Dq, A), self.parameter_init)
        if t in ("hybrid", "location"):
            k = 200 if t == "hybrid" else 201
            v["filter"] = tn((k, 1, 10), self.parameter_init if t == "hybrid" else 0.1)
            v["W_filter/weights"] = tn((10, A), self.parameter_init)
            v["W_filter/biases"] = np.zeros(A, np.float32)
        if t in ("bahdanau_content", "location", "hybrid", "luong_concat"):
            lim = np.sqrt(6.0 / (A + 1))
            v["v_a"] = rng.uniform(-lim, lim, A).astype(np.float32)
        self.variables = {k: torch.tensor(a, device=device) for k, a in v.items()}
        return self.variables

    # ------------------------------------------------------------- hoisted keys
    def precompute_keys(self, encoder_outputs):
        """Once per batch: the part of the energy that does not depend on the decoder state."""
        t = self.attention_type
        B, T, E = encoder_outputs.shape
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        enc2d = encoder_outputs.reshape(B * T, E)
        v = self.variables
        if t in ("bahdanau_content", "hybrid", "dot_product"):
            k = ops.gemm(enc2d, v["W_keys/weights"], False, False, v.get("W_keys/biases"), prec)
        elif t == "luong_general":
            k = ops.gemm(enc2d, v["W_keys/weights"], False, False, None, prec)
        elif t == "luong_concat":
            k = ops.gemm(enc2d, v["W_concat/weights"][:E], False, False, None, prec)
        elif t == "luong_dot":
            k = enc2d
        else:                         # location: no content term
            k = None
        self._keys = None if k is None else k.view(B, T, -1)
        return self._keys

    # ------------------------------------------------------------------ step
    def __call__(self, encoder_outputs, decoder_output, encoder_outputs_length, attention_weights,
                 out=None):
        """out: optional dict of preallocated cuda tensors {'alpha','context','energy','q'} the
        step writes into (the training decoder keeps them for the backward pass)."""
        lib = _lib.load()
        out = out or {}
        t = self.attention_type
        v = self.variables
        B, T, E = encoder_outputs.shape
        if self._keys is None and t != "location":
            self.precompute_keys(encoder_outputs)
        if t == "luong_dot" and E != decoder_output.shape[-1]:
            raise ValueError("encoder_num_units and decoder_num_units must be the same size.")
        if t in ("bahdanau_content", "location", "hybrid", "dot_product"):
            q = ops.gemm(decoder_output, v["W_query/weights"], out=out.get("q"))
        elif t == "luong_concat":
            q = ops.gemm(decoder_output, v["W_concat/weights"][E:], out=out.get("q"))
        else:
            q = decoder_output.contiguous()
        mode = 1 if t in ("dot_product", "luong_dot", "luong_general") else 0
        A = q.shape[-1]
        loc = t in ("hybrid", "location")
        alpha = out.get("alpha")
        if alpha is None:
            alpha = torch.empty((B, T), dtype=torch.float32, device=encoder_outputs.device)
        ctx = out.get("context")
        if ctx is None:
            ctx = torch.empty((B, E), dtype=torch.float32, device=encoder_outputs.device)
        energy = out.get("energy")
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else C.c_void_p(0)
        filt = v["filter"].reshape(-1, 10).contiguous() if loc else None
        rc = lib.b2_attention_step_forward(
            mode, p(encoder_outputs.contiguous()), p(self._keys), p(q),
            p(attention_weights.contiguous()) if loc else C.c_void_p(0),
            p(encoder_outputs_length), p(filt), filt.shape[0] if loc else 0,
            p(v["W_filter/weights"]) if loc else C.c_void_p(0),
            p(v["W_filter/biases"]) if loc else C.c_void_p(0),
            p(v.get("v_a")), B, T, E, A, float(self.sharpening_factor), int(bool(self.sigmoid_smoothing)),
            p(alpha), p(ctx), p(energy), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "b2_attention_step_forward")
        return alpha, ctx

    # -------------------------------------------------------------- backward
    def query_is_projected(self):
        return self.attention_type in ("bahdanau_content", "location", "hybrid", "dot_product", "luong_concat")

    def backward_step(self, encoder_outputs, q, alpha, energy, encoder_outputs_length, dctx, d_keys, dq,
                      grads, dq_accumulate=False):
        """Sequential part of the step's backward: dq [B,A] is written, d_keys / v_a / W_filter
        bias gradients accumulate.  ``grads``: dict with the same keys as ``variables``."""
        lib = _lib.load()
        t, v = self.attention_type, self.variables
        B, T, E = encoder_outputs.shape
        mode = 1 if t in ("dot_product", "luong_dot", "luong_general") else 0
        loc = t in ("hybrid", "location")
        A = q.shape[-1]
        nbytes = lib.b2_attention_step_backward_workspace_bytes(B, T)
        ws = ops.workspace("attn_bwd", nbytes, encoder_outputs.device)
        p = ops._ptr
        rc = lib.b2_attention_step_backward(
            mode, p(encoder_outputs), p(self._keys), p(q), p(alpha), p(energy), p(encoder_outputs_length),
            p(v["W_filter/biases"]) if loc else p(None), p(v.get("v_a")), B, T, E, A,
            float(self.sharpening_factor), int(bool(self.sigmoid_smoothing)), p(dctx), p(d_keys), p(dq),
            int(dq_accumulate), p(grads.get("v_a")), p(grads["W_filter/biases"]) if loc else p(None), p(ws), nbytes, ops._stream())
        _lib.check(rc, "b2_attention_step_backward")

    def backward_keys(self, encoder_outputs, d_keys, d_enc, grads):
        """After the loop: push the accumulated d_keys [B,T,A] through the hoisted key projection.
        d_enc [B,T,E] accumulates."""
        t, v = self.attention_type, self.variables
        B, T, E = encoder_outputs.shape
        if t in ("location", "luong_dot") or d_keys is None:
            return                                      # no projection (luong_dot: d_keys aliases d_enc)
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        enc2d, dk2d = encoder_outputs.reshape(B * T, E), d_keys.reshape(B * T, -1)
        if t == "luong_concat":
            w, gw = v["W_concat/weights"][:E], grads["W_concat/weights"][:E]
        else:
            w, gw = v["W_keys/weights"], grads["W_keys/weights"]
        ops.gemm(enc2d, dk2d, True, False, None, prec, out=gw, beta=1.0)
        if "W_keys/biases" in v and t != "luong_concat":
            ops.colsum(dk2d, out=grads["W_keys/biases"], accumulate=True)
        ops.gemm(dk2d, w, False, True, None, prec, out=d_enc.reshape(B * T, E), beta=1.0)
