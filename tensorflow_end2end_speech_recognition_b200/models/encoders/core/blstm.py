"""Bidirectional LSTM encoder -- host mirror of ``models/encoders/core/blstm.py``.

Same constructor and call signature as the reference class (blstm.py:13-121):
``BLSTMEncoder(num_units, num_proj, num_layers, lstm_impl, use_peephole,
parameter_init, clip_activation, time_major)`` and
``enc(inputs[B,T,D], inputs_seq_len[B], keep_prob, is_training) -> (outputs,
final_state)``.  Arithmetic: ``b2_blstm_layer_forward/backward`` (CUDA, sm_100a).
"""
import numpy as np
import torch

from .... import ops

LSTM_IMPLS = ("BasicLSTMCell", "LSTMCell", "LSTMBlockCell")


class BLSTMEncoder(object):
    def __init__(self, num_units, num_proj, num_layers, lstm_impl, use_peephole, parameter_init,
                 clip_activation, time_major=True, name="lstm_encoder", precision="fp32",
                 tf_version="1.2.0"):
        assert num_proj != 0
        self.num_units = num_units
        self.num_proj = num_proj if lstm_impl == "LSTMCell" else None       # blstm.py:49-52
        self.num_layers = num_layers
        self.lstm_impl = lstm_impl
        self.use_peephole = use_peephole
        self.parameter_init = parameter_init
        self.clip_activation = clip_activation
        self.time_major = time_major
        self.name = name
        self.precision = precision
        if lstm_impl in ("LSTMBlockFusedCell", "CudnnLSTM"):
            raise NotImplementedError("%s is 'under implementation' in the reference "
                                      "(blstm.py:335-473)" % lstm_impl)
        if lstm_impl not in LSTM_IMPLS:
            raise IndexError('lstm_impl is "BasicLSTMCell" or "LSTMCell" or ' +
                             '"LSTMBlockCell" or "LSTMBlockFusedCell" or ' + '"CudnnLSTM".')
        # BasicLSTMCell has no peephole / clip (blstm.py:148-157); the block cell only
        # receives clip_cell when tf.__version__ == '1.3.0' (blstm.py:286-305)
        if lstm_impl == "BasicLSTMCell":
            self._peephole, self._clip = False, None
        elif lstm_impl == "LSTMBlockCell":
            self._peephole = bool(use_peephole)
            self._clip = clip_activation if tf_version == "1.3.0" else None
        else:
            self._peephole, self._clip = bool(use_peephole), clip_activation
        self._saved = None
        self.num_layers_sub = None       # MultitaskBLSTMEncoder: the layer whose output feeds the sub task

    # ------------------------------------------------------------ variables
    def create_variables(self, input_size, rng):
        """-> ordered list of (tf_name, numpy array); U(-init, init) kernels and peepholes
        from the scope initializer (blstm.py:79-80,283-284), zero biases."""
        out = []
        d_in = input_size
        H = self.num_units
        Hout = self.num_proj or H                       # LSTMCell(num_proj): the recurrent / emitted width
        for i_layer in range(1, self.num_layers + 1):
            for d in ("fw", "bw"):
                scope = "blstm_hidden%d/%s/lstm_cell/" % (i_layer, d)
                a = self.parameter_init
                out.append((scope + "kernel", rng.uniform(-a, a, (d_in + Hout, 4 * H)).astype(np.float32)))
                out.append((scope + "bias", np.zeros(4 * H, np.float32)))
                if self._peephole:
                    for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
                        out.append((scope + k, rng.uniform(-a, a, H).astype(np.float32)))
                if self.num_proj:
                    out.append((scope + "projection/kernel",
                                rng.uniform(-a, a, (H, self.num_proj)).astype(np.float32)))
            d_in = 2 * Hout
        return out

    @property
    def output_size(self):
        return 2 * (self.num_proj or self.num_units)

    def _layer_params(self, variables, i_layer, d):
        scope = "blstm_hidden%d/%s/lstm_cell/" % (i_layer, d)
        p = {"kernel": variables[scope + "kernel"], "bias": variables[scope + "bias"]}
        if self._peephole:
            for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
                p[k] = variables[scope + k]
        if self.num_proj:
            p["projection"] = variables[scope + "projection/kernel"]
        return p

    # -------------------------------------------------------------- forward
    def __call__(self, inputs, inputs_seq_len, keep_prob, is_training, variables=None,
                 dropout_seed=0):
        """inputs [B,T,D] cuda f32 -> (outputs [T,B,2H] if time_major else [B,T,2H], final_state)."""
        assert variables is not None, "BLSTMEncoder needs the model's variable dict"
        B, T, D = inputs.shape
        x = ops.transpose_01(inputs)                       # blstm.py:279
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        saved = []
        fs = None
        x_lp = 0
        self.sub_outputs = self.sub_final_state = self.sub_output_lp = None
        for i_layer in range(1, self.num_layers + 1):
            desc = ops.lstm_desc(T, B, x.shape[2], self.num_units, use_peephole=self._peephole,
                                 forget_bias=1.0, cell_clip=self._clip, keep_prob=float(keep_prob),
                                 dropout_seed=dropout_seed * 131 + i_layer, precision=prec,
                                 need_backward=is_training, num_proj=self.num_proj)
            pf = self._layer_params(variables, i_layer, "fw")
            pb = self._layer_params(variables, i_layer, "bw")
            is_sub = self.num_layers_sub is not None and i_layer == self.num_layers_sub
            y, fs, reserve = ops.blstm_layer_forward(desc, x, inputs_seq_len, pf, pb,
                                                     want_final_state=(i_layer == self.num_layers or is_sub),
                                                     x_lp=x_lp)
            saved.append((desc, x, reserve, i_layer, x_lp))
            x = y
            x_lp = ops.reserve_y_lp(desc, reserve)   # bf16 shadow feeds the next layer's GEMM
            if is_sub:                               # blstm.py:325-327: outputs_sub / final_state_sub of that layer
                self.sub_outputs = y
                self.sub_final_state = ((fs[0], fs[1]), (fs[2], fs[3])) if fs is not None else None
                self.sub_output_lp = (x_lp, y.shape[2]) if x_lp else None
        self._saved = (saved, inputs_seq_len)
        # (device pointer, row stride) of the bf16 copy of the time-major output the last layer wrote, or None
        self.output_lp = (x_lp, x.shape[2]) if (x_lp and self.time_major) else None
        self._keepalive = reserve          # the shadow lives in the last layer's reserve
        outputs = x if self.time_major else ops.transpose_01(x)
        final_state = None
        if fs is not None:
            final_state = ((fs[0], fs[1]), (fs[2], fs[3]))   # (fw(c,h), bw(c,h))
        return outputs, final_state

    # ------------------------------------------------------------- backward
    def backward(self, d_outputs, variables, grads, need_dx=False, on_layer_done=None,
                 d_final_state=None, saved=None, d_inject=None):
        """d_outputs [T,B,2H] (time-major).  Accumulates into ``grads`` (same keys as
        ``variables``); calls ``on_layer_done(i_layer)`` when a layer's gradients are final.
        d_final_state [4,B,H]: gradient of the returned (fw(c,h), bw(c,h)) of the last layer.
        d_inject {layer: [T,B,2H]}: extra gradient of that layer's OUTPUT (heads tapped below the top)."""
        own = saved is None
        saved, seq_len = self._saved if own else saved
        dy = d_outputs
        # DropoutWrapper backward between two stacked layers that both run the tcgen05 path: layer l+1 applies layer l's
        # mask in the store of the dX GEMM it computes anyway, layer l then skips its own mask pass over dy (a full
        # read + write of [T,B,2H] per layer).  Not across a layer whose output also feeds a tapped head (d_inject:
        # that gradient is added unmasked).
        by_layer = {s[3]: s[0] for s in saved}
        fuse_below = {}
        for i_layer, desc in by_layer.items():
            lower = by_layer.get(i_layer - 1)
            fuse_below[i_layer] = bool(
                lower is not None and 0.0 < lower.keep_prob < 1.0 and not (d_inject and (i_layer - 1) in d_inject)
                and ops.blstm_layer_path(desc) == 1 and ops.blstm_layer_path(lower) == 1)
        for desc, x, reserve, i_layer, x_lp in reversed(saved):
            if fuse_below.get(i_layer) or fuse_below.get(i_layer + 1):
                lower = by_layer.get(i_layer - 1)
                desc = ops.lstm_desc_with(
                    desc, dy_premasked=int(bool(fuse_below.get(i_layer + 1))),
                    dx_keep_prob=float(lower.keep_prob) if fuse_below.get(i_layer) else 0.0,
                    dx_dropout_seed=int(lower.dropout_seed) if fuse_below.get(i_layer) else 0)
            if d_inject and i_layer in d_inject:      # gradient of a head tapped at this layer's output (sub task)
                dy = ops.add_(dy.contiguous(), d_inject[i_layer])
            pf = self._layer_params(variables, i_layer, "fw")
            pb = self._layer_params(variables, i_layer, "bw")
            gf = self._layer_params(grads, i_layer, "fw")
            gb = self._layer_params(grads, i_layer, "bw")
            dy = ops.blstm_layer_backward(desc, x, seq_len, pf, pb, dy, reserve, gf, gb,
                                          need_dx=(i_layer > 1 or need_dx), x_lp=x_lp,
                                          d_final_state=d_final_state if i_layer == self.num_layers else None)
            if on_layer_done is not None:
                on_layer_done(i_layer)
        ops.blstm_backward_join()      # side-stream weight-gradient GEMMs -> gradients final
        if own:
            self._saved = None
        return dy
