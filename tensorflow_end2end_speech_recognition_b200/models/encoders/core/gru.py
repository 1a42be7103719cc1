"""GRU encoders -- host mirror of ``models/encoders/core/gru.py``: ``GRUEncoder`` (gru.py:9-79, MultiRNNCell of
``tf.contrib.rnn.GRUCell`` under ``dynamic_rnn``) and ``BGRUEncoder`` (gru.py:82-160, one
``bidirectional_dynamic_rnn`` per layer).  Same constructor keywords (``num_units, num_layers, parameter_init,
time_major``) and ``enc(inputs[B,T,D], inputs_seq_len, keep_prob, is_training) -> (outputs, final_state)``; TF
variable names (``.../gru_cell/gates/kernel`` ...; the gates bias starts at 1 as in GRUCell).
Arithmetic: ``b2_bgru_layer_forward/backward`` (csrc/gru.cu); the unidirectional encoder binds an all-zero idle second
direction (its state, output and gradients are exactly zero)."""
import numpy as np
import torch

from .... import ops

_KEYS = ("gates/kernel", "gates/bias", "candidate/kernel", "candidate/bias")


class BGRUEncoder(object):
    bidirectional = True

    def __init__(self, num_units, num_layers, parameter_init, time_major=False, name="bgru_encoder", precision="fp32"):
        self.num_units, self.num_layers, self.parameter_init = num_units, num_layers, parameter_init
        self.time_major, self.name, self.precision = time_major, name, precision      # fp32 arithmetic either way
        self._saved = None
        self._idle = {}
        self.output_lp = None

    # ------------------------------------------------------------ variables
    def _scope(self, i_layer, d):
        if self.bidirectional:
            return "bgru_hidden%d/%s/gru_cell/" % (i_layer, d)
        return "multi_gru/multi_rnn_cell/cell_%d/gru_cell/" % (i_layer - 1)

    def create_variables(self, input_size, rng):
        out, d_in, H, a = [], input_size, self.num_units, self.parameter_init
        for i_layer in range(1, self.num_layers + 1):
            for d in (("fw", "bw") if self.bidirectional else ("fw",)):
                scope = self._scope(i_layer, d)
                out.append((scope + "gates/kernel", rng.uniform(-a, a, (d_in + H, 2 * H)).astype(np.float32)))
                out.append((scope + "gates/bias", np.ones(2 * H, np.float32)))
                out.append((scope + "candidate/kernel", rng.uniform(-a, a, (d_in + H, H)).astype(np.float32)))
                out.append((scope + "candidate/bias", np.zeros(H, np.float32)))
            d_in = self.output_size
        return out

    @property
    def output_size(self):
        return (2 if self.bidirectional else 1) * self.num_units

    def _layer_params(self, variables, i_layer, d):
        scope = self._scope(i_layer, d)
        return {k: variables[scope + k] for k in _KEYS}

    def _idle_direction(self, like):
        key = (tuple(like["gates/kernel"].shape), like["gates/kernel"].device)
        if key not in self._idle:
            self._idle[key] = ({k: torch.zeros_like(v) for k, v in like.items()},
                               {k: torch.zeros_like(v) for k, v in like.items()})
        return self._idle[key]

    # -------------------------------------------------------------- forward
    def __call__(self, inputs, inputs_seq_len, keep_prob, is_training, variables=None, dropout_seed=0):
        assert variables is not None, "GRU encoders need the model's variable dict"
        B, T, _ = inputs.shape
        H = self.num_units
        x = ops.transpose_01(inputs)
        saved, states = [], []
        for i_layer in range(1, self.num_layers + 1):
            desc = ops.gru_desc(T, B, x.shape[2], H, keep_prob=float(keep_prob),
                                dropout_seed=dropout_seed * 131 + i_layer, need_backward=is_training)
            pf = self._layer_params(variables, i_layer, "fw")
            pb = self._layer_params(variables, i_layer, "bw") if self.bidirectional else self._idle_direction(pf)[0]
            y, fs, reserve = ops.bgru_layer_forward(desc, x, inputs_seq_len, pf, pb, want_final_state=True)
            saved.append((desc, x, reserve, i_layer))
            x = y if self.bidirectional else y[:, :, :H].contiguous()
            states.append((fs[0], fs[1]) if self.bidirectional else fs[0])
        self._saved = (saved, inputs_seq_len)
        outputs = x if self.time_major else ops.transpose_01(x)
        # bidirectional_dynamic_rnn: (h_fw, h_bw) of the last layer; MultiRNNCell: the state of every layer
        return outputs, (states[-1] if self.bidirectional else tuple(states))

    # ------------------------------------------------------------- backward
    def backward(self, d_outputs, variables, grads, need_dx=False, on_layer_done=None, d_final_state=None,
                 saved=None):
        assert d_final_state is None, "GRU encoders: no bridge gradient path"
        own = saved is None
        saved, seq_len = self._saved if own else saved
        dy = d_outputs
        for desc, x, reserve, i_layer in reversed(saved):
            pf, gf = self._layer_params(variables, i_layer, "fw"), self._layer_params(grads, i_layer, "fw")
            if self.bidirectional:
                pb, gb = self._layer_params(variables, i_layer, "bw"), self._layer_params(grads, i_layer, "bw")
            else:
                pb, gb = self._idle_direction(pf)
                dy = torch.cat([dy, torch.zeros_like(dy)], dim=2)
            dy = ops.bgru_layer_backward(desc, x, seq_len, pf, pb, dy, reserve, gf, gb,
                                         need_dx=(i_layer > 1 or need_dx))
            if on_layer_done is not None:
                on_layer_done(i_layer)
        if own:
            self._saved = None
        return dy


class GRUEncoder(BGRUEncoder):
    bidirectional = False

    def __init__(self, num_units, num_layers, parameter_init, time_major=False, name="gru_encoder", precision="fp32"):
        super(GRUEncoder, self).__init__(num_units, num_layers, parameter_init, time_major=time_major, name=name,
                                         precision=precision)
