"""Joint CTC-attention model -- host mirror of ``models/attention/joint_ctc_attention.py``
(class ``JointCTCAttention``).

``compute_loss(inputs, labels, ctc_labels, inputs_seq_len, labels_seq_len, keep_prob_encoder,
keep_prob_decoder, keep_prob_embedding) -> (total_loss, logits, ctc_logits,
decoder_outputs_train, decoder_outputs_infer)`` with
``total = lambda * mean(ctc_loss) + (1 - lambda) * sequence_loss (+ L2)`` (:268-322).

The CTC head (``ctc_logits`` :182-235) is one fully connected layer on the shared encoder
states.  The reference flattens the *batch-major* states and reshapes the result as
``[T, B, C]`` (:223-226), which scrambles frames for B > 1: the default here
is the intended time-major head, ``faithful_ctc_reshape=True`` reproduces the reference.
As in the reference (:124-138) the constructor ignores the caller's
``clip_activation_decoder``, ``weight_decay``, ``time_major``, ``sharpening_factor`` and
``logits_temperature`` unless ``honour_all_kwargs=True``.
"""
import numpy as np
import torch

from ... import ops
from ...compat import graph as _graph
from ...compat.graph import graph_op
from ..ctc.ctc import _truncated_normal
from ...utils.io.labels.sparsetensor import SparseTensorValue, sparse_to_label_lists
from .attention_seq2seq import AttentionSeq2Seq


class JointCTCAttention(AttentionSeq2Seq):
    def __init__(self, input_size, encoder_type, encoder_num_units, encoder_num_layers,
                 encoder_num_proj, attention_type, attention_dim, decoder_type, decoder_num_units,
                 decoder_num_layers, embedding_dim, lambda_weight, num_classes, sos_index, eos_index,
                 max_decode_length, lstm_impl="LSTMBlockCell", use_peephole=True, splice=1,
                 parameter_init=0.1, clip_grad_norm=5.0, clip_activation_encoder=50,
                 clip_activation_decoder=50, weight_decay=0.0, time_major=True,
                 sharpening_factor=1.0, logits_temperature=1.0, name="joint_ctc_attention",
                 honour_all_kwargs=False, faithful_ctc_reshape=False, **b200_kwargs):
        self.ctc_num_classes = num_classes + 1                  # + blank (:140)
        self.lambda_weight = lambda_weight
        self.faithful_ctc_reshape = faithful_ctc_reshape
        self.ctc_labels_pl_list = []
        if not honour_all_kwargs:                               # joint_ctc_attention.py:124-138
            clip_activation_decoder, weight_decay, time_major = 50, 0.0, True
            sharpening_factor, logits_temperature = 1.0, 1.0
        super(JointCTCAttention, self).__init__(
            input_size=input_size, encoder_type=encoder_type, encoder_num_units=encoder_num_units,
            encoder_num_layers=encoder_num_layers, encoder_num_proj=encoder_num_proj,
            attention_type=attention_type, attention_dim=attention_dim, decoder_type=decoder_type,
            decoder_num_units=decoder_num_units, decoder_num_layers=decoder_num_layers,
            embedding_dim=embedding_dim, num_classes=num_classes, sos_index=sos_index,
            eos_index=eos_index, max_decode_length=max_decode_length, lstm_impl=lstm_impl,
            use_peephole=use_peephole, splice=splice, parameter_init=parameter_init,
            clip_grad_norm=clip_grad_norm, clip_activation_encoder=clip_activation_encoder,
            clip_activation_decoder=clip_activation_decoder, weight_decay=weight_decay,
            time_major=time_major, sharpening_factor=sharpening_factor,
            logits_temperature=logits_temperature, name=name, **b200_kwargs)

    def _extra_variables(self, rng):
        E = 2 * self.encoder_num_units
        return [("ctc_output/weights", _truncated_normal(rng, (E, self.ctc_num_classes), self.parameter_init)),
                ("ctc_output/biases", np.zeros(self.ctc_num_classes, np.float32))]

    def create_placeholders(self):
        super(JointCTCAttention, self).create_placeholders()
        P = _graph.Placeholder
        self.ctc_labels_pl_list.append(_graph.SparseTensor(P("int64"), P("int32"), P("int64")))

    def ctc_logits(self, encoder_outputs):
        """encoder_outputs [B,T,2H] batch-major -> logits [T,B,ctc_num_classes]  (:182-235)"""
        B, T, E = encoder_outputs.shape
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        w, b = self.variables["ctc_output/weights"], self.variables["ctc_output/biases"]
        if self.faithful_ctc_reshape:
            rows = encoder_outputs.reshape(B * T, E)            # b-major rows, read as [T,B] below
        else:
            rows = self._enc_tm.view(T * B, E)
        self._ctc_rows = rows
        return ops.gemm(rows, w, False, False, b, prec).view(T, B, self.ctc_num_classes)

    @graph_op(n_out=5, name="compute_loss")
    def compute_loss(self, inputs, labels, ctc_labels, inputs_seq_len, labels_seq_len,
                     keep_prob_encoder, keep_prob_decoder, keep_prob_embedding, scope=None,
                     is_training=True):
        inputs = self._dev(inputs, torch.float32)
        labels = self._dev(labels, torch.int32)
        inputs_seq_len = self._dev(inputs_seq_len, torch.int32)
        labels_seq_len = self._dev(labels_seq_len, torch.int32)
        B = inputs.shape[0]
        if isinstance(ctc_labels, SparseTensorValue) or (
                isinstance(ctc_labels, (list, tuple)) and len(ctc_labels) == 3 and
                getattr(ctc_labels[0], "ndim", 0) == 2):
            ctc_lists = sparse_to_label_lists(ctc_labels, B)
        else:
            ctc_lists = [list(l) for l in ctc_labels]
        ops.check_labels(ctc_lists, self.ctc_num_classes, self.ctc_num_classes - 1, what="ctc_labels")
        logits, out_train, out_infer, enc_bm = self._build(
            inputs, labels, inputs_seq_len, labels_seq_len, keep_prob_encoder, keep_prob_decoder,
            keep_prob_embedding, is_training)
        lam = float(self.lambda_weight)
        seq_part, dlogits = self._sequence_loss(logits, labels, labels_seq_len, 1.0 - lam, is_training)
        ctc_logits = self.ctc_logits(enc_bm)
        # ignore_longer_outputs_than_inputs=False (:315): a label longer than its input is an error
        lens_host = inputs_seq_len.cpu().numpy()
        for b, l in enumerate(ctc_lists):
            if len(l) > int(lens_host[b]):
                raise RuntimeError("Not enough time for target transition sequence "
                                   "(required: %d, available: %d)" % (len(l), int(lens_host[b])))
        flat, offs, lmax = ops.pack_labels(ctc_lists)
        d_flat = torch.as_tensor(flat).to(self.device, non_blocking=True)
        d_offs = torch.as_tensor(offs).to(self.device, non_blocking=True)
        losses, d_ctc = ops.ctc_loss_grad(ctc_logits, d_flat, d_offs, inputs_seq_len, lmax,
                                          blank=self.ctc_num_classes - 1, ignore_longer=False,
                                          grad_scale=lam / B, need_grad=is_training)
        self.ctc_losses = losses
        self.ctc_loss = losses.mean()
        total_loss = seq_part + lam * self.ctc_loss
        total_loss = self._add_weight_decay(total_loss)
        self._ctx = {"dlogits": dlogits, "shape": tuple(inputs.shape), "d_ctc": d_ctc} if is_training else None
        return total_loss, logits, ctc_logits, out_train, out_infer

    def _backward_extra(self, d_enc_tm):
        """CTC head backward; adds its d(encoder states) to the attention path's."""
        d_ctc = self._ctx["d_ctc"]
        T, B, C = d_ctc.shape
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        dl2d = d_ctc.view(T * B, C)
        ops.gemm(self._ctc_rows, dl2d, True, False, None, prec, out=self.grads["ctc_output/weights"], beta=1.0)
        ops.colsum(dl2d, out=self.grads["ctc_output/biases"], accumulate=True)
        w = self.variables["ctc_output/weights"]
        if self.faithful_ctc_reshape:
            # rows were the batch-major states: gradient lands in batch-major order
            d_bm = ops.gemm(dl2d, w, False, True, None, prec).view(B, T, -1)
            d_extra = ops.transpose_01(d_bm)
            ops.axpy_multi(ops.TensorList([d_extra]), ops.TensorList([d_enc_tm]), 1.0)
        else:
            ops.gemm(dl2d, w, False, True, None, prec, out=d_enc_tm.view(T * B, -1), beta=1.0)
        return d_enc_tm
