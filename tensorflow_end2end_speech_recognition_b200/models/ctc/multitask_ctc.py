"""Multi-task (hierarchical) CTC model -- host mirror of ``models/ctc/multitask_ctc.py`` (class
``MultitaskCTC``, :14-420): a BLSTM stack with the main CTC head on the top layer and a second CTC head on
layer ``num_layers_sub``; ``total_loss = main_task_weight * ctc_main + (1 - main_task_weight) * ctc_sub``
(:268-296, both with ``ignore_longer_outputs_than_inputs=False``).

Same constructor keywords and method names as the reference:
    compute_loss(inputs, labels_main, labels_sub, inputs_seq_len, keep_prob, scope=None)
        -> (total_loss, logits_main [T,B,C_main], logits_sub [T,B,C_sub])
    decoder(logits_main, logits_sub, inputs_seq_len, beam_width=1) -> (decode_main, decode_sub)
    posteriors(logits_main, logits_sub), compute_ler(decode_main, decode_sub, labels_main, labels_sub)
Arithmetic: the CUDA kernels behind ``CTC`` (BLSTM layers, ``b2_ctc_loss_grad`` twice, GEMM heads).
"""
import numpy as np
import torch

from ... import ops
from ...compat import graph as _graph
from ...compat.graph import graph_op
from ..encoders.load_encoder import load
from ..model_base import ModelBase
from .ctc import CTC, _truncated_normal, label_lists_from, ler_from_lists


class MultitaskCTC(CTC):
    def __init__(self, encoder_type, input_size, num_units, num_layers_main, num_layers_sub,
                 num_classes_main, num_classes_sub, main_task_weight, lstm_impl="LSTMBlockCell",
                 use_peephole=True, splice=1, parameter_init=0.1, clip_grad_norm=None, clip_activation=None,
                 num_proj=None, weight_decay=0.0, bottleneck_dim=None, time_major=True,
                 precision="fp32", device=None, seed=1):
        ModelBase.__init__(self)
        assert splice % 2 == 1, "splice must be the odd number"
        if clip_grad_norm is not None:
            assert float(clip_grad_norm) > 0, "clip_grad_norm must be larger than 0."
        assert float(weight_decay) >= 0, "weight_decay must not be a negative value."
        if float(main_task_weight) < 0 or float(main_task_weight) > 1:
            raise ValueError("Set main_task_weight between 0 to 1.")                 # multitask_ctc.py:90-91
        if encoder_type not in ("multitask_blstm", "multitask_lstm"):
            raise NotImplementedError("encoder_type %r: 'multitask_blstm' and 'multitask_lstm' are built" % (encoder_type,))
        self.encoder_type, self.input_size, self.splice, self.num_stack = encoder_type, input_size, splice, 1
        self.num_units = num_units
        self.num_proj = int(num_proj) if num_proj not in (None, 0, "0") else None
        self.num_layers = num_layers_main
        self.num_layers_sub = num_layers_sub
        self.bottleneck_dim = bottleneck_dim
        self.num_classes = num_classes_main + 1             # + blank
        self.num_classes_sub = num_classes_sub + 1
        self.main_task_weight = float(main_task_weight)
        self.sub_task_weight = 1.0 - self.main_task_weight
        self.lstm_impl, self.use_peephole, self.parameter_init = lstm_impl, use_peephole, parameter_init
        self.clip_grad_norm, self.clip_activation, self.weight_decay = clip_grad_norm, clip_activation, weight_decay
        self.summaries_train, self.summaries_dev = [], []
        self.inputs_pl_list, self.labels_pl_list, self.labels_sub_pl_list = [], [], []
        self.inputs_seq_len_pl_list, self.keep_prob_pl_list = [], []
        self.time_major = time_major
        self.name = encoder_type + "_ctc"
        self.precision = precision
        self.device = torch.device(device if device is not None else "cuda:0")
        self.encoder = load(encoder_type)(
            num_units=num_units, num_proj=self.num_proj, num_layers_main=num_layers_main,
            num_layers_sub=num_layers_sub, lstm_impl=lstm_impl, use_peephole=use_peephole,
            parameter_init=parameter_init, clip_activation=clip_activation, time_major=True, precision=precision)
        rng = np.random.RandomState(seed)
        named = self.encoder.create_variables(input_size * splice, rng)
        out_in = self.encoder.output_size
        named.append(("output_sub/weights", _truncated_normal(rng, (out_in, self.num_classes_sub), parameter_init)))
        named.append(("output_sub/biases", np.zeros(self.num_classes_sub, np.float32)))
        main_in = out_in
        if self.bottleneck_dim not in (None, 0):
            named.append(("bottleneck/weights", _truncated_normal(rng, (out_in, int(self.bottleneck_dim)), parameter_init)))
            named.append(("bottleneck/biases", np.zeros(int(self.bottleneck_dim), np.float32)))
            main_in = int(self.bottleneck_dim)
        named.append(("output_main/weights", _truncated_normal(rng, (main_in, self.num_classes), parameter_init)))
        named.append(("output_main/biases", np.zeros(self.num_classes, np.float32)))
        self._allocate_variables(named, self.device)
        self._step = 0
        self._ctx = None
        decay = [v for v in self._variables if "bias" not in v.name.lower()]
        self._decay_params = ops.TensorList([v.tensor for v in decay])
        self._decay_grads = ops.TensorList([v.grad for v in decay])

    def create_placeholders(self):
        """(multitask_ctc.py:193-211)"""
        super(MultitaskCTC, self).create_placeholders()
        self.labels_sub_pl_list.append(_graph.SparseTensor(_graph.Placeholder("int64", name="indices_sub"),
                                                           _graph.Placeholder("int32", name="values_sub"),
                                                           _graph.Placeholder("int64", name="shape_sub")))

    # ----------------------------------------------------------------- model
    def _build(self, inputs, inputs_seq_len, keep_prob, is_training=True):
        """-> (logits_main [T,B,C_main], logits_sub [T,B,C_sub])   (multitask_ctc.py:109-191)"""
        B, T, _ = inputs.shape
        self._step += 1
        enc, _, enc_sub, _ = self.encoder(inputs, inputs_seq_len, keep_prob, is_training,
                                          variables=self.variables, dropout_seed=self._step)
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        lp_main = self.encoder.output_lp if prec == ops.PREC_BF16 else None
        lp_sub = self.encoder.sub_output_lp if prec == ops.PREC_BF16 else None
        feat_sub = enc_sub.view(T * B, -1)
        logits_sub = ops.gemm(feat_sub, self.variables["output_sub/weights"], False, False,
                              self.variables["output_sub/biases"], prec, a_lp=lp_sub)
        feat = enc.view(T * B, -1)
        bneck = None
        head_lp = lp_main
        if self.bottleneck_dim not in (None, 0):
            feat = ops.gemm(feat, self.variables["bottleneck/weights"], False, False,
                            self.variables["bottleneck/biases"], prec, a_lp=lp_main)
            ops.relu_dropout_(feat, float(keep_prob), self._step * 7919 + 5)
            bneck, head_lp = (feat, float(keep_prob)), None
        logits = ops.gemm(feat, self.variables["output_main/weights"], False, False,
                          self.variables["output_main/biases"], prec, a_lp=head_lp)
        self._state = {"enc": enc, "enc_lp": lp_main, "enc_sub": enc_sub, "sub_lp": lp_sub, "head_in": feat,
                       "head_lp": head_lp, "bneck": bneck, "saved": self.encoder._saved}
        return logits.view(T, B, self.num_classes), logits_sub.view(T, B, self.num_classes_sub)

    def _ctc(self, logits, labels, seq_len_dev, seq_len_host, num_classes, weight, B, is_training, what):
        lists = label_lists_from(labels, B)
        ops.check_labels(lists, num_classes, num_classes - 1, what=what)
        for b, l in enumerate(lists):          # ignore_longer_outputs_than_inputs=False (:272,:285)
            if len(l) > int(seq_len_host[b]):
                raise RuntimeError("Not enough time for target transition sequence "
                                   "(required: %d, available: %d)" % (len(l), int(seq_len_host[b])))
        flat, offs, lmax = ops.pack_labels(lists)
        d_flat = torch.as_tensor(flat).to(self.device, non_blocking=True)
        d_offs = torch.as_tensor(offs).to(self.device, non_blocking=True)
        losses, dlogits = ops.ctc_loss_grad(logits, d_flat, d_offs, seq_len_dev, lmax, blank=num_classes - 1,
                                            ignore_longer=False, grad_scale=weight / B, need_grad=is_training)
        return losses.mean(), dlogits

    @graph_op(n_out=3, name="compute_loss")
    def compute_loss(self, inputs, labels_main, labels_sub, inputs_seq_len, keep_prob, scope=None,
                     is_training=True):
        inputs, inputs_seq_len = self._to_device(inputs, inputs_seq_len)
        B = inputs.shape[0]
        logits_main, logits_sub = self._build(inputs, inputs_seq_len, keep_prob, is_training)
        lens_host = inputs_seq_len.cpu().numpy()
        l_main, d_main = self._ctc(logits_main, labels_main, inputs_seq_len, lens_host, self.num_classes,
                                   self.main_task_weight, B, is_training, "labels_main")
        l_sub, d_sub = self._ctc(logits_sub, labels_sub, inputs_seq_len, lens_host, self.num_classes_sub,
                                 self.sub_task_weight, B, is_training, "labels_sub")
        self.ctc_loss_main, self.ctc_loss_sub = l_main, l_sub
        total_loss = self.main_task_weight * l_main + self.sub_task_weight * l_sub
        if self.weight_decay > 0:
            sq = ops.clip_by_norm_multi(self._decay_params, 3.0e38)
            total_loss = total_loss + 0.5 * float(self.weight_decay) * sq.sum()
        self._ctx = (d_main, d_sub, tuple(inputs.shape), self._state) if is_training else None
        total_loss._b2_ctx = self._ctx
        return total_loss, logits_main, logits_sub

    def _backward(self, ctx=None, flat=None, grads=None):
        ctx = ctx if ctx is not None else self._ctx
        assert ctx is not None, "train() needs a preceding compute_loss(is_training=True)"
        d_main, d_sub, (B, T, _), st = ctx
        flat = self.flat_grads if flat is None else flat
        grads = self.grads if grads is None else grads
        flat.zero_()
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        # sub head (tapped at layer num_layers_sub)
        ds2d = d_sub.view(T * B, self.num_classes_sub)
        sub2d = st["enc_sub"].view(T * B, -1)
        ops.gemm(sub2d, ds2d, True, False, None, prec, out=grads["output_sub/weights"], beta=1.0, a_lp=st["sub_lp"])
        ops.colsum(ds2d, out=grads["output_sub/biases"], accumulate=True)
        denc_sub = ops.gemm(ds2d, self.variables["output_sub/weights"], False, True, None, prec)
        # main head (+ bottleneck)
        dm2d = d_main.view(T * B, self.num_classes)
        ops.gemm(st["head_in"], dm2d, True, False, None, prec, out=grads["output_main/weights"], beta=1.0,
                 a_lp=st["head_lp"])
        ops.colsum(dm2d, out=grads["output_main/biases"], accumulate=True)
        denc = ops.gemm(dm2d, self.variables["output_main/weights"], False, True, None, prec)
        if st["bneck"] is not None:
            feat, kp = st["bneck"]
            dz = ops.relu_dropout_backward(denc, feat, kp)
            ops.gemm(st["enc"].view(T * B, -1), dz, True, False, None, prec, out=grads["bottleneck/weights"],
                     beta=1.0, a_lp=st["enc_lp"])
            ops.colsum(dz, out=grads["bottleneck/biases"], accumulate=True)
            denc = ops.gemm(dz, self.variables["bottleneck/weights"], False, True, None, prec)
        self.encoder.backward(denc.view(T, B, -1), self.variables, grads, d_outputs_sub=denc_sub.view(T, B, -1),
                              saved=st["saved"])
        if self.weight_decay > 0:
            decay_grads = self._decay_grads if grads is self.grads else \
                ops.TensorList([grads[v.name] for v in self._variables if "bias" not in v.name.lower()])
            ops.axpy_multi(self._decay_params, decay_grads, float(self.weight_decay))
        if ctx is self._ctx:
            self._ctx = None

    # ---------------------------------------------------------------- decode
    def decoder(self, logits_main, logits_sub, inputs_seq_len, beam_width=1):
        """-> (decode_op_main, decode_op_sub)   (multitask_ctc.py:322-354)"""
        return (CTC.decoder(self, logits_main, inputs_seq_len, beam_width),
                CTC.decoder(self, logits_sub, inputs_seq_len, beam_width))

    def posteriors(self, logits_main, logits_sub):
        """(multitask_ctc.py:356-382)"""
        if _graph.is_handle(logits_main) or _graph.is_handle(logits_sub):
            return (_graph.Op(lambda l: ops.softmax_rows(ops.transpose_01(l).view(-1, self.num_classes)),
                              (logits_main,), {}, name="posteriors_main"),
                    _graph.Op(lambda l: ops.softmax_rows(ops.transpose_01(l).view(-1, self.num_classes_sub)),
                              (logits_sub,), {}, name="posteriors_sub"))
        return (ops.softmax_rows(ops.transpose_01(logits_main).view(-1, self.num_classes)),
                ops.softmax_rows(ops.transpose_01(logits_sub).view(-1, self.num_classes_sub)))

    def compute_ler(self, decode_op_main, decode_op_sub, labels_main, labels_sub):
        """-> (ler_main, ler_sub)   (multitask_ctc.py:384-420)"""
        return (CTC.compute_ler(self, decode_op_main, labels_main),
                CTC.compute_ler(self, decode_op_sub, labels_sub))
