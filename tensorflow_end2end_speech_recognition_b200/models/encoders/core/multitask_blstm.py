"""Multi-task bidirectional LSTM encoder -- host mirror of ``models/encoders/core/multitask_blstm.py``
(class ``MultitaskBLSTMEncoder``, :13-125): the BLSTM stack of ``blstm.py`` with a second output tapped after
layer ``num_layers_sub`` (``blstm.py:325-331``: ``outputs_sub`` / ``final_state_sub``).  Same constructor keywords,
``enc(inputs, inputs_seq_len, keep_prob, is_training) -> (outputs, final_state, outputs_sub, final_state_sub)``.
Arithmetic: the same ``b2_blstm_layer_forward/backward`` launches as ``BLSTMEncoder``; the sub task's gradient enters
the BPTT of layer ``num_layers_sub`` through ``backward(d_outputs_sub=...)``.
"""
from .blstm import BLSTMEncoder


class MultitaskBLSTMEncoder(BLSTMEncoder):
    def __init__(self, num_units, num_proj, num_layers_main, num_layers_sub, lstm_impl, use_peephole,
                 parameter_init, clip_activation, time_major=False, name="multitask_blstm_encoder",
                 precision="fp32", tf_version="1.2.0"):
        super(MultitaskBLSTMEncoder, self).__init__(
            num_units=num_units, num_proj=num_proj, num_layers=num_layers_main, lstm_impl=lstm_impl,
            use_peephole=use_peephole, parameter_init=parameter_init, clip_activation=clip_activation,
            time_major=time_major, name=name, precision=precision, tf_version=tf_version)
        if num_layers_sub < 1 or num_layers_main < num_layers_sub:
            raise ValueError("Set num_layers_sub between 1 to num_layers_main.")     # multitask_ctc.py:235-239
        self.num_layers_main = num_layers_main
        self.num_layers_sub = num_layers_sub

    def __call__(self, inputs, inputs_seq_len, keep_prob, is_training=True, variables=None, dropout_seed=0):
        outputs, final_state = super(MultitaskBLSTMEncoder, self).__call__(
            inputs, inputs_seq_len, keep_prob, is_training, variables=variables, dropout_seed=dropout_seed)
        outputs_sub = self.sub_outputs
        if not self.time_major:
            from .... import ops
            outputs_sub = ops.transpose_01(outputs_sub)
        return outputs, final_state, outputs_sub, self.sub_final_state

    def backward(self, d_outputs, variables, grads, d_outputs_sub=None, **kw):
        """d_outputs / d_outputs_sub: time-major [T,B,2H] gradients of the two returned output tensors"""
        inject = {self.num_layers_sub: d_outputs_sub} if d_outputs_sub is not None else None
        if inject and self.num_layers_sub == self.num_layers:      # both heads on the top layer
            from .... import ops
            d_outputs = ops.add_(d_outputs.contiguous(), d_outputs_sub)
            inject = None
        return super(MultitaskBLSTMEncoder, self).backward(d_outputs, variables, grads, d_inject=inject, **kw)
