"""Unidirectional LSTM encoder -- host mirror of ``models/encoders/core/lstm.py``.

Same constructor and call signature as the reference class (lstm.py:13-118):
``LSTMEncoder(num_units, num_proj, num_layers, lstm_impl, use_peephole, parameter_init,
clip_activation, time_major)`` and ``enc(inputs[B,T,D], inputs_seq_len[B], keep_prob,
is_training) -> (outputs [.., num_units], final_state)``; variables are named as
``tf.contrib.rnn.MultiRNNCell`` names them (``multi_lstm/multi_rnn_cell/cell_<i>/lstm_cell/...``,
lstm.py:127-166).

Arithmetic: the BLSTM layer kernels (``b2_blstm_layer_forward/backward``) with the second direction
idle -- its weights are a zero block, so its cell state, output and input gradient are exactly
zero and the forward direction's half of the [T,B,2H] output IS the unidirectional layer.  That
costs one idle direction of compute; a one-direction launch of the recurrence kernels is not built
(widening row, not the headline path).
"""
import numpy as np
import torch

from .... import ops
from .blstm import BLSTMEncoder


class LSTMEncoder(BLSTMEncoder):
    def __init__(self, num_units, num_proj, num_layers, lstm_impl, use_peephole, parameter_init,
                 clip_activation, time_major=False, name="lstm_encoder", precision="fp32",
                 tf_version="1.2.0"):
        super(LSTMEncoder, self).__init__(num_units, num_proj, num_layers, lstm_impl, use_peephole,
                                          parameter_init, clip_activation, time_major=time_major, name=name,
                                          precision=precision, tf_version=tf_version)
        if self.num_proj:
            raise NotImplementedError("LSTMEncoder with num_proj: the projected recurrence exists for the "
                                      "bidirectional encoder only")
        self._idle = {}          # (D, device) -> (zero parameters, scratch gradients) of the idle direction

    # ------------------------------------------------------------ variables
    @staticmethod
    def _scope(i_layer):
        return "multi_lstm/multi_rnn_cell/cell_%d/lstm_cell/" % (i_layer - 1)

    def create_variables(self, input_size, rng):
        out = []
        d_in, H, a = input_size, self.num_units, self.parameter_init
        for i_layer in range(1, self.num_layers + 1):
            scope = self._scope(i_layer)
            out.append((scope + "kernel", rng.uniform(-a, a, (d_in + H, 4 * H)).astype(np.float32)))
            out.append((scope + "bias", np.zeros(4 * H, np.float32)))
            if self._peephole:
                for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
                    out.append((scope + k, rng.uniform(-a, a, H).astype(np.float32)))
            d_in = H
        return out

    @property
    def output_size(self):
        return self.num_units

    def _layer_params(self, variables, i_layer, d="fw"):
        scope = self._scope(i_layer)
        p = {"kernel": variables[scope + "kernel"], "bias": variables[scope + "bias"]}
        if self._peephole:
            for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
                p[k] = variables[scope + k]
        return p

    def _idle_direction(self, like):
        key = (tuple(like["kernel"].shape), like["kernel"].device)
        if key not in self._idle:
            zeros = {k: torch.zeros_like(v) for k, v in like.items()}
            scratch = {k: torch.zeros_like(v) for k, v in like.items()}
            self._idle[key] = (zeros, scratch)
        return self._idle[key]

    # -------------------------------------------------------------- forward
    def __call__(self, inputs, inputs_seq_len, keep_prob, is_training, variables=None, dropout_seed=0):
        """inputs [B,T,D] cuda f32 -> (outputs [T,B,H] if time_major else [B,T,H], final_state:
        ((c, h) of every layer), as MultiRNNCell's dynamic_rnn returns it)."""
        assert variables is not None, "LSTMEncoder needs the model's variable dict"
        B, T, D = inputs.shape
        H = self.num_units
        x = ops.transpose_01(inputs)
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        saved, states = [], []
        for i_layer in range(1, self.num_layers + 1):
            desc = ops.lstm_desc(T, B, x.shape[2], H, use_peephole=self._peephole, forget_bias=1.0,
                                 cell_clip=self._clip, keep_prob=float(keep_prob),
                                 dropout_seed=dropout_seed * 131 + i_layer, precision=prec,
                                 need_backward=is_training)
            pf = self._layer_params(variables, i_layer)
            idle, _ = self._idle_direction(pf)
            y2, fs, reserve = ops.blstm_layer_forward(desc, x, inputs_seq_len, pf, idle, want_final_state=True)
            saved.append((desc, x, reserve, i_layer))
            x = y2[:, :, :H].contiguous()
            states.append((fs[0], fs[1]))
            if self.num_layers_sub is not None and i_layer == self.num_layers_sub:     # lstm.py: outputs_sub
                self.sub_outputs, self.sub_final_state = x, tuple(states)
        self._saved = (saved, inputs_seq_len)
        self.output_lp = None
        self.sub_output_lp = None
        if self.num_layers_sub is None:
            self.sub_outputs = self.sub_final_state = None
        outputs = x if self.time_major else ops.transpose_01(x)
        return outputs, tuple(states)

    # ------------------------------------------------------------- backward
    def backward(self, d_outputs, variables, grads, need_dx=False, on_layer_done=None, d_final_state=None,
                 saved=None, d_inject=None):
        """d_outputs [T,B,H] (time-major).  Accumulates into ``grads``; returns d(inputs) [T,B,D] or None."""
        assert d_final_state is None, "LSTMEncoder: no bridge gradient path"
        own = saved is None
        saved, seq_len = self._saved if own else saved
        dy = d_outputs
        for desc, x, reserve, i_layer in reversed(saved):
            if d_inject and i_layer in d_inject:      # gradient of a head tapped at this layer's output (sub task)
                dy = ops.add_(dy.contiguous(), d_inject[i_layer])
            pf = self._layer_params(variables, i_layer)
            gf = self._layer_params(grads, i_layer)
            idle, scratch = self._idle_direction(pf)
            dy2 = torch.cat([dy, torch.zeros_like(dy)], dim=2)     # the idle direction receives no gradient
            dy = ops.blstm_layer_backward(desc, x, seq_len, pf, idle, dy2, reserve, gf, scratch,
                                          need_dx=(i_layer > 1 or need_dx))
            if on_layer_done is not None:
                on_layer_done(i_layer)
        ops.blstm_backward_join()
        if own:
            self._saved = None
        return dy


class MultitaskLSTMEncoder(LSTMEncoder):
    """``models/encoders/core/multitask_lstm.py``: the unidirectional stack with a second output tapped after layer
    ``num_layers_sub``; ``enc(...) -> (outputs, final_state, outputs_sub, final_state_sub)``."""

    def __init__(self, num_units, num_proj, num_layers_main, num_layers_sub, lstm_impl, use_peephole,
                 parameter_init, clip_activation, time_major=False, name="multitask_lstm_encoder",
                 precision="fp32", tf_version="1.2.0"):
        super(MultitaskLSTMEncoder, self).__init__(num_units, num_proj, num_layers_main, lstm_impl, use_peephole,
                                                   parameter_init, clip_activation, time_major=time_major, name=name,
                                                   precision=precision, tf_version=tf_version)
        if num_layers_sub < 1 or num_layers_main < num_layers_sub:
            raise ValueError("Set num_layers_sub between 1 to num_layers_main.")
        self.num_layers_main, self.num_layers_sub = num_layers_main, num_layers_sub

    def __call__(self, inputs, inputs_seq_len, keep_prob, is_training=True, variables=None, dropout_seed=0):
        outputs, final_state = super(MultitaskLSTMEncoder, self).__call__(
            inputs, inputs_seq_len, keep_prob, is_training, variables=variables, dropout_seed=dropout_seed)
        outputs_sub = self.sub_outputs if self.time_major else ops.transpose_01(self.sub_outputs)
        return outputs, final_state, outputs_sub, self.sub_final_state

    def backward(self, d_outputs, variables, grads, d_outputs_sub=None, **kw):
        inject = {self.num_layers_sub: d_outputs_sub} if d_outputs_sub is not None else None
        if inject and self.num_layers_sub == self.num_layers:
            d_outputs = ops.add_(d_outputs.contiguous(), d_outputs_sub)
            inject = None
        return super(MultitaskLSTMEncoder, self).backward(d_outputs, variables, grads, d_inject=inject, **kw)
