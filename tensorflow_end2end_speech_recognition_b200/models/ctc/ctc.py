"""CTC model -- host mirror of ``models/ctc/ctc.py`` (class ``CTC``, :16-398).

Same constructor signature and method names as the reference; handles are
evaluated eagerly on CUDA tensors instead of being TF graph ops:

    model = CTC(encoder_type='blstm', input_size=120, num_units=256, num_layers=2,
                num_classes=61, ...)
    loss, logits = model.compute_loss(inputs, labels_st, inputs_seq_len, keep_prob)
    model.train(loss, optimizer='rmsprop', learning_rate=1e-3)
    decoded = model.decoder(logits, inputs_seq_len, beam_width=1)   # SparseTensorValue
    ler = model.compute_ler(decoded, labels_st)
"""
import numpy as np
import torch

from ... import ops
from ...compat import graph as _graph
from ...compat.graph import graph_op
from ...utils.io.labels.sparsetensor import SparseTensorValue, sparse_to_label_lists
from ..encoders.load_encoder import load
from ..model_base import ModelBase


def _truncated_normal(rng, shape, stddev):
    """tf.truncated_normal: re-draw until within 2 sigma (ctc.py:221-223)."""
    x = rng.normal(0.0, stddev, size=shape)
    bad = np.abs(x) > 2 * stddev
    while bad.any():
        x[bad] = rng.normal(0.0, stddev, size=int(bad.sum()))
        bad = np.abs(x) > 2 * stddev
    return x.astype(np.float32)


class CTC(ModelBase):
    def __init__(self, encoder_type, input_size, num_units, num_layers, num_classes,
                 lstm_impl="LSTMBlockCell", use_peephole=True, splice=1, num_stack=1,
                 parameter_init=0.1, clip_grad_norm=None, clip_activation=None, num_proj=None,
                 weight_decay=0.0, bottleneck_dim=None, time_major=True,
                 precision="fp32", device=None, seed=1, strict_input_size=False):
        super(CTC, self).__init__()
        # The reference asserts input_size % 3 == 0 (ctc.py:79) although the BLSTM path never
        # uses it; BASELINE's 80-d features need the relaxed form.
        if strict_input_size:
            assert input_size % 3 == 0, \
                "input_size must be divisible by 3 (+ delta, acceleration coefficients)."
        assert splice % 2 == 1, "splice must be the odd number"
        if clip_grad_norm is not None:
            assert float(clip_grad_norm) > 0, "clip_grad_norm must be larger than 0."
        assert float(weight_decay) >= 0, "weight_decay must not be a negative value."

        self.encoder_type = encoder_type
        self.input_size = input_size
        self.splice = splice
        self.num_stack = num_stack
        self.num_units = num_units
        # the reference evaluates int(num_proj) before the None check (ctc.py:93)
        self.num_proj = int(num_proj) if num_proj not in (None, 0, "0") else None
        self.num_layers = num_layers
        self.bottleneck_dim = bottleneck_dim
        self.num_classes = num_classes + 1      # + blank (ctc.py:101)
        self.lstm_impl = lstm_impl
        self.use_peephole = use_peephole
        self.parameter_init = parameter_init
        self.clip_grad_norm = clip_grad_norm
        self.clip_activation = clip_activation
        self.weight_decay = weight_decay
        self.summaries_train, self.summaries_dev = [], []
        self.inputs_pl_list, self.labels_pl_list = [], []
        self.inputs_seq_len_pl_list, self.keep_prob_pl_list = [], []
        self.time_major = time_major
        self.name = encoder_type + "_ctc"
        self.precision = precision
        self.device = torch.device(device if device is not None else "cuda:0")

        if encoder_type in ["blstm", "lstm"]:
            self.encoder = load(encoder_type)(
                num_units=num_units, num_proj=self.num_proj, num_layers=num_layers,
                lstm_impl=lstm_impl, use_peephole=use_peephole, parameter_init=parameter_init,
                clip_activation=clip_activation, time_major=True, precision=precision)
        elif encoder_type in ["bgru", "gru"]:                       # ctc.py: GRU encoders take no cell options
            self.encoder = load(encoder_type)(num_units=num_units, num_layers=num_layers,
                                              parameter_init=parameter_init, time_major=True, precision=precision)
        elif encoder_type in ["vgg_blstm", "vgg_lstm"]:
            self.encoder = load(encoder_type)(
                input_size=input_size, splice=splice, num_stack=num_stack, num_units=num_units,
                num_proj=self.num_proj, num_layers=num_layers, lstm_impl=lstm_impl,
                use_peephole=use_peephole, parameter_init=parameter_init,
                clip_activation=clip_activation, time_major=True, precision=precision)
        else:
            raise NotImplementedError(
                "encoder_type %r: 'blstm', 'lstm', 'bgru', 'gru', 'vgg_blstm' and 'vgg_lstm' are built on the B200 kernels" %
                (encoder_type,))

        rng = np.random.RandomState(seed)
        named = self.encoder.create_variables(input_size * num_stack * splice, rng)
        out_in = self.encoder.output_size if hasattr(self.encoder, "output_size") else 2 * num_units
        if self.bottleneck_dim not in (None, 0):                       # ctc.py:200-209
            named.append(("bottleneck/weights", _truncated_normal(rng, (out_in, int(self.bottleneck_dim)),
                                                                  parameter_init)))
            named.append(("bottleneck/biases", np.zeros(int(self.bottleneck_dim), np.float32)))
            out_in = int(self.bottleneck_dim)
        named.append(("output/weights", _truncated_normal(rng, (out_in, self.num_classes), parameter_init)))
        named.append(("output/biases", np.zeros(self.num_classes, np.float32)))
        self._allocate_variables(named, self.device)
        self._step = 0
        self._ctx = None
        # weight decay covers every variable whose name has no 'bias' (ctc.py:283-285): kernels,
        # peepholes and the output weights
        decay = [v for v in self._variables if "bias" not in v.name.lower()]
        self._decay_params = ops.TensorList([v.tensor for v in decay])
        self._decay_grads = ops.TensorList([v.grad for v in decay])

    # ----------------------------------------------------------------- feeds
    def create_placeholders(self):
        """One placeholder set per tower (ctc.py:240-254).  Passing these handles to compute_loss /
        decoder / ... yields lazy ops for ``compat.tf.Session.run``; passing arrays evaluates eagerly."""
        from ...compat import tf as _tf
        _tf.register_model(self)
        self.inputs_pl_list.append(_graph.Placeholder("float32", [None, None, self.input_size], "input"))
        self.labels_pl_list.append(_graph.SparseTensor(_graph.Placeholder("int64", name="indices"),
                                                       _graph.Placeholder("int32", name="values"),
                                                       _graph.Placeholder("int64", name="shape")))
        self.inputs_seq_len_pl_list.append(_graph.Placeholder("int32", [None], "inputs_seq_len"))
        self.keep_prob_pl_list.append(_graph.Placeholder("float32", name="keep_prob"))

    def _to_device(self, inputs, inputs_seq_len):
        if not torch.is_tensor(inputs):
            inputs = torch.as_tensor(np.ascontiguousarray(inputs, dtype=np.float32))
        if not inputs.is_cuda:
            inputs = inputs.to(self.device, non_blocking=True)
        if not torch.is_tensor(inputs_seq_len):
            inputs_seq_len = torch.as_tensor(np.asarray(inputs_seq_len, dtype=np.int32))
        if not inputs_seq_len.is_cuda:
            inputs_seq_len = inputs_seq_len.to(self.device, non_blocking=True)
        return inputs.float(), inputs_seq_len.int()

    # ----------------------------------------------------------------- model
    def _build(self, inputs, inputs_seq_len, keep_prob, is_training):
        """encoder -> output FC -> logits [T,B,num_classes] (ctc.py:175-238)."""
        B, T, _ = inputs.shape
        self._step += 1
        enc, final_state = self.encoder(inputs, inputs_seq_len, keep_prob, is_training,
                                        variables=self.variables, dropout_seed=self._step)
        self.encoder_outputs = enc
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        feat = enc.view(T * B, -1)
        # the last BLSTM layer already wrote a bf16 copy of its output: the output-layer GEMMs read that
        # instead of re-casting 64000 x 1024 floats (twice per step)
        feat_lp = getattr(self.encoder, "output_lp", None) if prec == ops.PREC_BF16 else None
        self._enc_lp = feat_lp
        self._bneck = None
        if self.bottleneck_dim not in (None, 0):
            # fully connected + ReLU, then dropout on the hidden-output connection (ctc.py:200-213)
            feat = ops.gemm(feat, self.variables["bottleneck/weights"], False, False,
                            self.variables["bottleneck/biases"], prec, a_lp=feat_lp)
            ops.relu_dropout_(feat, float(keep_prob), self._step * 7919 + 5)
            self._bneck = (feat, float(keep_prob))
            feat_lp = None
        self._head_in = feat
        self._head_lp = feat_lp
        logits2d = ops.gemm(feat, self.variables["output/weights"], False, False,
                            self.variables["output/biases"], prec, a_lp=feat_lp)
        return logits2d.view(T, B, self.num_classes)

    def compute_loss(self, inputs, labels, inputs_seq_len, keep_prob, scope=None,
                     softmax_temperature=1, is_training=True):
        """-> (total_loss 0-d cuda tensor, logits [T,B,C])   (ctc.py:256-323).

        labels: the SparseTensor triple (indices, values, dense_shape) of
        ``list2sparsetensor`` or a list of label sequences.

        Graph mode (any argument a placeholder): two ops, as in the reference's graph -- ``logits`` depends
        only on (inputs, inputs_seq_len, keep_prob), the loss on (logits, labels, inputs_seq_len) -- so that
        ``sess.run(decode_op)`` / ``sess.run(posteriors_op)`` need no labels in the feed
        (examples/timit/metrics/ctc.py:72-81) and do not run the CTC loss."""
        if any(_graph.is_handle(a) for a in (inputs, labels, inputs_seq_len, keep_prob)):
            logits_op = _graph.Op(self._eval_logits, (inputs, inputs_seq_len, keep_prob, is_training), {},
                                  name="logits")
            loss_op = _graph.Op(self._eval_loss, (logits_op, labels, inputs_seq_len, softmax_temperature,
                                                  is_training), {}, name="compute_loss")
            return loss_op, logits_op
        logits = self._eval_logits(inputs, inputs_seq_len, keep_prob, is_training)
        return self._eval_loss(logits, labels, inputs_seq_len, softmax_temperature, is_training), logits

    def _eval_logits(self, inputs, inputs_seq_len, keep_prob, is_training=True):
        inputs, inputs_seq_len = self._to_device(inputs, inputs_seq_len)
        logits = self._build(inputs, inputs_seq_len, keep_prob, is_training)
        # everything the backward pass of THIS forward needs travels with the logits, so that several towers
        # (examples/librispeech/training/train_ctc.py:82-147) can be in flight on one model object
        logits._b2_state = {"head_in": self._head_in, "bneck": self._bneck, "enc": self.encoder_outputs,
                            "saved": getattr(self.encoder, "_saved", None),
                            "head_lp": self._head_lp, "enc_lp": self._enc_lp}
        return logits

    def _eval_loss(self, logits, labels, inputs_seq_len, softmax_temperature=1, is_training=True):
        _, inputs_seq_len = self._to_device(logits, inputs_seq_len)
        T, B, _ = logits.shape
        label_lists = label_lists_from(labels, B)
        ops.check_labels(label_lists, self.num_classes, self.num_classes - 1)
        flat, offs, lmax = ops.pack_labels(label_lists)
        d_flat = torch.as_tensor(flat).to(self.device, non_blocking=True)
        d_offs = torch.as_tensor(offs).to(self.device, non_blocking=True)
        scaled = logits if softmax_temperature == 1 else logits / float(softmax_temperature)
        # grad of reduce_mean(ctc_losses) (ctc.py:298) wrt logits comes out of the same launch
        losses, dlogits = ops.ctc_loss_grad(scaled, d_flat, d_offs, inputs_seq_len, lmax,
                                            blank=self.num_classes - 1, ignore_longer=True,
                                            grad_scale=1.0 / (B * float(softmax_temperature)),
                                            need_grad=is_training)
        self.ctc_losses = losses
        total_loss = losses.mean()
        if self.weight_decay > 0:
            # weight_decay * sum_{non-bias} l2_loss(w), l2_loss = sum(w^2)/2   (ctc.py:280-286)
            sq = ops.clip_by_norm_multi(self._decay_params, 3.0e38)      # norms^2, no scaling
            total_loss = total_loss + 0.5 * float(self.weight_decay) * sq.sum()
        self._ctx = (dlogits, (B, T, None), getattr(logits, "_b2_state", None)) if is_training else None
        total_loss._b2_ctx = self._ctx
        return total_loss

    def _backward(self, ctx=None, flat=None, grads=None):
        """Gradients of a compute_loss into ``flat`` / ``grads`` (default: the model's own flat_grads views).
        ``ctx``: the context of the loss to differentiate (default: the last compute_loss)."""
        ctx = ctx if ctx is not None else self._ctx
        assert ctx is not None, "train() needs a preceding compute_loss(is_training=True)"
        dlogits, (B, T, _), state = ctx
        flat = self.flat_grads if flat is None else flat
        grads = self.grads if grads is None else grads
        head_in, bneck, enc_out = self._head_in, self._bneck, self.encoder_outputs
        head_lp, enc_lp = getattr(self, "_head_lp", None), getattr(self, "_enc_lp", None)
        saved = None
        if state is not None:
            head_in, bneck, enc_out, saved = state["head_in"], state["bneck"], state["enc"], state["saved"]
            head_lp, enc_lp = state["head_lp"], state["enc_lp"]
        flat.zero_()
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        enc2d = enc_out.view(T * B, -1)
        dl2d = dlogits.view(T * B, self.num_classes)
        ops.gemm(head_in, dl2d, True, False, None, prec, out=grads["output/weights"], beta=1.0, a_lp=head_lp)
        ops.colsum(dl2d, out=grads["output/biases"], accumulate=True)
        denc = ops.gemm(dl2d, self.variables["output/weights"], False, True, None, prec)
        if bneck is not None:
            feat, kp = bneck
            dz = ops.relu_dropout_backward(denc, feat, kp)
            ops.gemm(enc2d, dz, True, False, None, prec, out=grads["bottleneck/weights"], beta=1.0, a_lp=enc_lp)
            ops.colsum(dz, out=grads["bottleneck/biases"], accumulate=True)
            denc = ops.gemm(dz, self.variables["bottleneck/weights"], False, True, None, prec)
        if getattr(self, "_on_heads_done", None) is not None:
            self._on_heads_done()
        self.encoder.backward(denc.view(T, B, -1), self.variables, grads, saved=saved,
                              on_layer_done=self._on_layer_done)
        if self.weight_decay > 0:
            decay_grads = self._decay_grads if grads is self.grads else \
                ops.TensorList([grads[v.name] for v in self._variables if "bias" not in v.name.lower()])
            ops.axpy_multi(self._decay_params, decay_grads, float(self.weight_decay))
        if ctx is self._ctx:
            self._ctx = None

    # ---------------------------------------------------------------- decode
    @graph_op(name="decoder")
    def decoder(self, logits, inputs_seq_len, beam_width=1, merge_repeated=True, impl="tf"):
        """-> SparseTensorValue(indices int64 [N,2], values int32 [N], dense_shape)  (ctc.py:325-352)

        beam_width == 1: ``tf.nn.ctc_greedy_decoder``; > 1: ``tf.nn.ctc_beam_search_decoder(beam_width,
        top_paths=1, merge_repeated=True)`` semantics (``b2_ctc_beam_decode_tf``) -- the op the reference model
        calls.  ``impl="numpy"`` selects the reference's OTHER decoder, the numpy prefix beam search of
        ``models/ctc/decoders/beam_search_decoder.py`` (no output merge, ``b2_ctc_beam_decode``)."""
        assert isinstance(beam_width, int), "beam_width must be integer."
        assert beam_width >= 1, "beam_width must be >= 1"
        _, inputs_seq_len = self._to_device(logits, inputs_seq_len)
        T, B, C = logits.shape
        if beam_width == 1:
            lab, n = ops.ctc_greedy_decode(logits, inputs_seq_len, blank=C - 1)
        elif impl == "numpy":
            lp = torch.log(ops.softmax_rows(ops.transpose_01(logits)))
            lab, n, _ = ops.ctc_beam_decode(lp, inputs_seq_len, beam_width, blank=C - 1)
        else:
            lab, n, _ = ops.ctc_beam_decode_tf(logits, inputs_seq_len, beam_width, blank=C - 1,
                                               merge_repeated=merge_repeated)
        lab, n = lab.cpu().numpy(), n.cpu().numpy()
        idx, val = [], []
        for b in range(B):
            for j in range(int(n[b])):
                idx.append((b, j))
                val.append(lab[b, j])
        maxlen = int(n.max()) if B else 0
        return SparseTensorValue(np.asarray(idx, np.int64).reshape(-1, 2), np.asarray(val, np.int32),
                                 np.asarray([B, maxlen], np.int64))

    @graph_op(name="posteriors")
    def posteriors(self, logits, blank_prior=1):
        """softmax over classes, batch-major [B*T, num_classes]  (ctc.py:354-380)"""
        lb = ops.transpose_01(logits)
        return ops.softmax_rows(lb.view(-1, self.num_classes))

    @graph_op(name="compute_ler")
    def compute_ler(self, decode_op, labels):
        """mean_b edit_distance(hyp_b, ref_b)/len(ref_b)  (ctc.py:382-398)"""
        B = int(decode_op.dense_shape[0])
        hyp = sparse_to_label_lists(decode_op, B)
        ref = label_lists_from(labels, B)
        return ler_from_lists(hyp, ref, self.device)


def label_lists_from(labels, B):
    """the three label formats every entry point accepts -> list of B python lists:
    the ``list2sparsetensor`` triple (indices [N,2], values [N], dense_shape [2]) -- recognised by
    ``indices.ndim == 2`` so that a batch of exactly three 1-D label arrays is not mistaken for it --,
    a SparseTensorValue, or a list of label sequences."""
    if isinstance(labels, SparseTensorValue):
        return sparse_to_label_lists(labels, B)
    if isinstance(labels, (list, tuple)) and len(labels) == 3 and getattr(labels[0], "ndim", 0) == 2:
        return sparse_to_label_lists(labels, B)
    return [list(l) for l in labels]


def ler_from_lists(hyp, ref, device):
    """mean_b edit_distance / len(ref_b), distances from b2_edit_distance (tf.edit_distance, normalize=True)"""
    d = ops.edit_distance(hyp, ref, device).astype(np.float64)
    n = np.asarray([len(r) for r in ref], np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.mean(d / n))


def _edit_distance(hyp, ref):
    """host form, kept for the CPU-side tests of the wire formats"""
    n, m = len(hyp), len(ref)
    d = list(range(m + 1))
    for i in range(1, n + 1):
        prev, d[0] = d[0], i
        for j in range(1, m + 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (hyp[i - 1] != ref[j - 1]))
            prev = cur
    return d[m]
