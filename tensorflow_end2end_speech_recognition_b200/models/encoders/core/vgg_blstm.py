"""VGG + bidirectional LSTM encoder -- host mirror of ``models/encoders/core/vgg_blstm.py``.

Same constructor and call signature as the reference class (vgg_blstm.py:40-75, :77-220):
``VGGBLSTMEncoder(input_size, splice, num_stack, num_units, num_proj, num_layers, lstm_impl,
use_peephole, parameter_init, clip_activation, time_major)`` and ``enc(inputs[B,T,D],
inputs_seq_len, keep_prob, is_training) -> (outputs, final_state)``.  The VGG front-end is
``b2_vgg_frontend_forward/backward`` (CUDA), its 256-d output feeds the BLSTM stack of
``core/blstm.py`` unchanged.
"""
import numpy as np
import torch

from .... import ops
from .blstm import BLSTMEncoder


def _truncated_normal(rng, shape, stddev):
    x = rng.normal(0.0, stddev, size=shape)
    bad = np.abs(x) > 2 * stddev
    while bad.any():
        x[bad] = rng.normal(0.0, stddev, size=int(bad.sum()))
        bad = np.abs(x) > 2 * stddev
    return x.astype(np.float32)


class VGGBLSTMEncoder(object):
    VGG_NAMES = ops.VGG_CONVS

    def __init__(self, input_size, splice, num_stack, num_units, num_proj, num_layers, lstm_impl,
                 use_peephole, parameter_init, clip_activation, time_major=False,
                 name="vgg_blstm_encoder", precision="fp32", tf_version="1.2.0"):
        assert num_proj != 0
        assert input_size % 3 == 0
        self.num_channels = input_size // 3
        self.splice, self.num_stack = splice, num_stack
        self.num_units, self.num_layers = num_units, num_layers
        self.lstm_impl, self.use_peephole = lstm_impl, use_peephole
        self.parameter_init, self.clip_activation = parameter_init, clip_activation
        self.time_major, self.name = time_major, name
        self.precision = precision
        self.blstm = BLSTMEncoder(num_units, num_proj, num_layers, lstm_impl, use_peephole, parameter_init,
                                  clip_activation, time_major=True, precision=precision, tf_version=tf_version)
        self.num_proj = self.blstm.num_proj
        self._saved = None

    @property
    def output_size(self):
        return self.blstm.output_size

    # ------------------------------------------------------------ variables
    def create_variables(self, input_size, rng):
        """-> ordered [(tf_name, numpy)]: conv filters/biases (cnn_util.py:67-70: truncated normal,
        zero bias), the 256-unit bridge (vgg_blstm.py:165-172), then the BLSTM stack on 256-d input."""
        W = self.splice * self.num_stack
        assert input_size == self.num_channels * W * 3
        out = []
        chans = (3, 64, 64, 128, 128)
        for i, n in enumerate(self.VGG_NAMES):
            out.append((n + "/weight", _truncated_normal(rng, (3, 3, chans[i], chans[i + 1]), self.parameter_init)))
            out.append((n + "/bias", np.zeros(chans[i + 1], np.float32)))
        h2, w2 = (self.num_channels + 1) // 2, (W + 1) // 2
        flat = ((h2 + 1) // 2) * ((w2 + 1) // 2) * 128
        out.append(("bridge/weights", _truncated_normal(rng, (flat, 256), self.parameter_init)))
        out.append(("bridge/biases", np.zeros(256, np.float32)))
        out += self.blstm.create_variables(256, rng)
        return out

    def _vgg_params(self, d):
        keys = [n + s for n in self.VGG_NAMES for s in ("/weight", "/bias")] + ["bridge/weights", "bridge/biases"]
        return {k: d[k] for k in keys}

    # -------------------------------------------------------------- forward
    def __call__(self, inputs, inputs_seq_len, keep_prob, is_training, variables=None, dropout_seed=0):
        assert variables is not None, "VGGBLSTMEncoder needs the model's variable dict"
        B, T, D = inputs.shape
        W = self.splice * self.num_stack
        assert D == self.num_channels * W * 3                          # vgg_blstm.py:106
        desc = ops.vgg_desc(B * T, self.num_channels, W, keep_prob=float(keep_prob),
                            dropout_seed=(dropout_seed * 977 + 13) * 8,
                            precision=ops.PREC_BF16 if getattr(self, "precision", "fp32") == "bf16" else ops.PREC_FP32)
        # [B,T,D] -> [B*T, num_channels, W, 3] is a pure reshape (:108-110)
        feat, reserve = ops.vgg_frontend_forward(desc, inputs.contiguous(), self._vgg_params(variables))
        outputs, final_state = self.blstm(feat.view(B, T, 256), inputs_seq_len, keep_prob, is_training,
                                          variables=variables, dropout_seed=dropout_seed)
        self.output_lp = getattr(self.blstm, "output_lp", None) if self.time_major else None
        # (front-end state, BLSTM state): one bundle, so a tower's backward can be handed its own
        self._saved = (desc, reserve, self.blstm._saved) if is_training else None
        if not self.time_major:
            outputs = ops.transpose_01(outputs)
        return outputs, final_state

    # ------------------------------------------------------------- backward
    def backward(self, d_outputs, variables, grads, need_dx=False, on_layer_done=None, d_final_state=None,
                 saved=None):
        """d_outputs time-major [T,B,2H]."""
        own = saved is None
        desc, reserve, blstm_saved = self._saved if own else saved
        d_feat_tm = self.blstm.backward(d_outputs, variables, grads, need_dx=True, on_layer_done=on_layer_done,
                                        d_final_state=d_final_state, saved=blstm_saved)   # [T,B,256]
        if own:
            self.blstm._saved = None
        d_feat = ops.transpose_01(d_feat_tm)                                     # [B,T,256]
        ops.vgg_frontend_backward(desc, self._vgg_params(variables), d_feat.view(-1, 256), reserve,
                                  self._vgg_params(grads))
        if own:
            self._saved = None
        return None
