"""Attention seq2seq model -- host mirror of ``models/attention/attention_seq2seq.py``
(class ``AttentionSeq2Seq``, :37-760).

Same constructor keywords and method names as the reference; handles are evaluated eagerly:

    model = AttentionSeq2Seq(input_size=120, encoder_type='blstm', encoder_num_units=256, ...)
    loss, logits, out_train, out_infer = model.compute_loss(
        inputs, labels, inputs_seq_len, labels_seq_len, kp_enc, kp_dec, kp_emb)
    model.train(loss, optimizer='adam', learning_rate=1e-3)
    ids_train, ids_infer = model.decode(out_train, out_infer)

``_build`` (:196-283): encoder -> bridge -> teacher-forced decoder (+ greedy inference
decoder, evaluated lazily: the reference only runs it when ``decoded_infer`` is fetched).
Deviation documented in DESIGN.md: GRU decoder not built (the reference's GRU branch references an
undefined attribute).
"""
from collections import namedtuple

import numpy as np
import torch

from ... import ops
from ...compat import graph as _graph
from ...compat.graph import graph_op
from ..encoders.load_encoder import load
from ..model_base import ModelBase
from .bridge import InitialStateBridge, LSTMStateTuple
from .decoders.attention_decoder import AttentionDecoder, AttentionDecoderOutput, LSTMBlockCell
from .decoders.attention_layer import AttentionLayer
from .decoders.helpers import GreedyEmbeddingHelper, TrainingHelper

EncoderOutput = namedtuple("EncoderOutput", ["outputs", "final_state", "seq_len"])

_ENC = "encoder/"
_BRIDGE = "decoder/bridge/"
_EMB = "decoder/output_embedding/W_embedding"
_CELL = "decoder/decoder_rnn_cell/lstm_cell/"
_ATT = "decoder/attention_decoder/attention_layer/"
_DEC = "decoder/attention_decoder/"


class _LazyDecoderOutput(object):
    """AttentionDecoderOutput of the inference decoder, computed on first access."""

    def __init__(self, fn):
        self._fn, self._val = fn, None

    def _get(self):
        if self._val is None:
            self._val = self._fn()
        return self._val

    def __getattr__(self, name):
        if name in AttentionDecoderOutput._fields:
            return getattr(self._get(), name)
        raise AttributeError(name)

    def __iter__(self):
        return iter(self._get())


class AttentionSeq2Seq(ModelBase):
    def __init__(self, input_size, encoder_type, encoder_num_units, encoder_num_layers,
                 encoder_num_proj, attention_type, attention_dim, decoder_type, decoder_num_units,
                 decoder_num_layers, embedding_dim, num_classes, sos_index, eos_index,
                 max_decode_length, lstm_impl="LSTMBlockCell", use_peephole=True, splice=1,
                 parameter_init=0.1, clip_grad_norm=5.0, clip_activation_encoder=50,
                 clip_activation_decoder=50, weight_decay=0.0, time_major=True,
                 sharpening_factor=1.0, logits_temperature=1.0, sigmoid_smoothing=False,
                 name="attention", precision="fp32", device=None, seed=1,
                 strict_input_size=False, tf_version="1.2.0", feed_previous_attention=False):
        super(AttentionSeq2Seq, self).__init__()
        if strict_input_size:                                   # attention_seq2seq.py:128 (see CTC)
            assert input_size % 3 == 0, \
                "input_size must be divisible by 3 (+ delta, double delta features)."
        assert splice % 2 == 1, "splice must be the odd number"
        assert clip_grad_norm > 0, "clip_grad_norm must be larger than 0."
        assert weight_decay >= 0, "weight_decay must not be a negative value."
        if encoder_type != "blstm":
            raise NotImplementedError("encoder_type %r: only 'blstm' is built" % (encoder_type,))
        if decoder_type != "lstm":
            if decoder_type == "gru":
                raise NotImplementedError("GRU decoder is not built (the reference's branch is broken, "
                                          "attention_seq2seq.py:365)")
            raise TypeError('decoder_type is "lstm" or "gru".')
        self.input_size, self.splice = input_size, splice
        self.encoder_type = encoder_type
        self.encoder_num_units, self.encoder_num_proj = encoder_num_units, encoder_num_proj
        self.encoder_num_layers = encoder_num_layers
        self.lstm_impl, self.use_peephole = lstm_impl, use_peephole
        self.attention_type, self.attention_dim = attention_type, attention_dim
        self.sharpening_factor, self.sigmoid_smoothing = sharpening_factor, sigmoid_smoothing
        self.decoder_type, self.decoder_num_units = decoder_type, decoder_num_units
        self.decdoder_num_layers = decoder_num_layers            # (sic) attention_seq2seq.py:158
        self.embedding_dim = embedding_dim
        self.num_classes = num_classes + 2                       # + <SOS>, <EOS>  (:160)
        self.sos_index, self.eos_index = sos_index, eos_index
        self.max_decode_length = max_decode_length
        self.logits_temperature = logits_temperature
        self.use_beam_search = False
        self.parameter_init = parameter_init
        self.clip_grad_norm = clip_grad_norm
        self.clip_activation_encoder = clip_activation_encoder
        self.clip_activation_decoder = clip_activation_decoder
        self.weight_decay = weight_decay
        self.time_major = time_major
        self.name = name
        self.precision = precision
        self.device = torch.device(device if device is not None else "cuda:0")
        self.summaries_train, self.summaries_dev = [], []
        for k in ("inputs", "labels", "inputs_seq_len", "labels_seq_len", "keep_prob_encoder",
                  "keep_prob_decoder", "keep_prob_embedding", "labels_st_true", "labels_st_pred"):
            setattr(self, k + "_pl_list", [])

        self.encoder = load(encoder_type)(
            num_units=encoder_num_units, num_proj=None, num_layers=encoder_num_layers,
            lstm_impl=lstm_impl, use_peephole=use_peephole, parameter_init=parameter_init,
            clip_activation=clip_activation_encoder, time_major=True, precision=precision,
            tf_version=tf_version)
        E = 2 * encoder_num_units
        self.attention_layer = AttentionLayer(attention_type, attention_dim, parameter_init,
                                              sharpening_factor, sigmoid_smoothing, precision=precision)
        cell = LSTMBlockCell(decoder_num_units, forget_bias=1.0,
                             clip_cell=clip_activation_decoder if tf_version == "1.3.0" else None,
                             use_peephole=use_peephole)
        self.decoder = AttentionDecoder(cell, parameter_init, max_decode_length, self.num_classes, None, None,
                                        self.attention_layer, time_major=False,
                                        feed_previous_attention=feed_previous_attention)
        self.bridge = InitialStateBridge(None, cell.state_size, parameter_init)

        rng = np.random.RandomState(seed)
        cpu = torch.device("cpu")
        named = [(_ENC + n, a) for n, a in self.encoder.create_variables(input_size * splice, rng)]
        named += [(_BRIDGE + k.split("/")[1], v.numpy()) for k, v in
                  self.bridge.create_variables(4 * encoder_num_units, rng, cpu).items()]
        named.append((_EMB, rng.uniform(-parameter_init, parameter_init,
                                        (self.num_classes, embedding_dim)).astype(np.float32)))
        dec_vars = self.decoder._create_variables(embedding_dim, E, rng, cpu)
        named += [(_CELL + k, v.numpy()) for k, v in self.decoder.cell_variables.items()]
        named += [(_ATT + k, v.numpy()) for k, v in
                  self.attention_layer.create_variables(E, decoder_num_units, rng, cpu).items()]
        named += [(_DEC + k, v.numpy()) for k, v in dec_vars.items()]
        named += self._extra_variables(rng)
        self._allocate_variables(named, self.device)
        # hand the sub-objects views into the flat buffers
        sub = lambda d, pre: {k[len(pre):]: v for k, v in d.items() if k.startswith(pre)}
        self._enc_vars, self._enc_grads = sub(self.variables, _ENC), sub(self.grads, _ENC)
        self.bridge.variables = {"bridge/" + k: v for k, v in sub(self.variables, _BRIDGE).items()}
        self._bridge_grads = {"bridge/" + k: v for k, v in sub(self.grads, _BRIDGE).items()}
        self.decoder.cell_variables, self._cell_grads = sub(self.variables, _CELL), sub(self.grads, _CELL)
        self.attention_layer.variables, self._att_grads = sub(self.variables, _ATT), sub(self.grads, _ATT)
        dec_keys = ("attentional_vector/weights", "output_layer/weights", "output_layer/biases")
        self.decoder.variables = {k: self.variables[_DEC + k] for k in dec_keys}
        self._dec_grads = {k: self.grads[_DEC + k] for k in dec_keys}
        decay = [v for v in self._variables if "bias" not in v.name.lower()]
        self._decay_params = ops.TensorList([v.tensor for v in decay])
        self._decay_grads = ops.TensorList([v.grad for v in decay])
        self._step = 0
        self._ctx = None

    def _extra_variables(self, rng):
        return []

    # ----------------------------------------------------------------- feeds
    def create_placeholders(self):
        """Graph-mode relic (attention_seq2seq.py:511-548); feeds go straight to compute_loss."""
        from ...compat import tf as _tf
        _tf.register_model(self)
        P, S = _graph.Placeholder, _graph.SparseTensor
        self.inputs_pl_list.append(P("float32", [None, None, self.input_size], "input"))
        self.labels_pl_list.append(P("int32", [None, None], "labels"))
        self.inputs_seq_len_pl_list.append(P("int32", [None], "inputs_seq_len"))
        self.labels_seq_len_pl_list.append(P("int32", [None], "labels_seq_len"))
        self.keep_prob_encoder_pl_list.append(P("float32", name="keep_prob_encoder"))
        self.keep_prob_decoder_pl_list.append(P("float32", name="keep_prob_decoder"))
        self.keep_prob_embedding_pl_list.append(P("float32", name="keep_prob_embedding"))
        self.labels_st_true_pl = S(P("int64"), P("int32"), P("int64"))
        self.labels_st_pred_pl = S(P("int64"), P("int32"), P("int64"))
        self.labels_st_true_pl_list.append(S(P("int64"), P("int32"), P("int64")))
        self.labels_st_pred_pl_list.append(S(P("int64"), P("int32"), P("int64")))

    def _dev(self, x, dtype):
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.ascontiguousarray(x))
        return x.to(self.device, non_blocking=True).to(dtype).contiguous()

    # ----------------------------------------------------------------- model
    def _encode(self, inputs, inputs_seq_len, keep_prob_encoder, is_training=True):
        """-> EncoderOutput(outputs [B,T,2H] batch-major, final_state, seq_len)  (:285-320)"""
        self._step += 1
        enc_tm, final_state = self.encoder(inputs, inputs_seq_len, keep_prob_encoder, is_training,
                                           variables=self._enc_vars, dropout_seed=self._step)
        self._enc_tm = enc_tm
        return EncoderOutput(ops.transpose_01(enc_tm), final_state, inputs_seq_len)

    def _build(self, inputs, labels, inputs_seq_len, labels_seq_len, keep_prob_encoder,
               keep_prob_decoder, keep_prob_embedding, is_training=True):
        enc = self._encode(inputs, inputs_seq_len, keep_prob_encoder, is_training)
        self.decoder.encoder_outputs = enc.outputs
        self.decoder.encoder_outputs_seq_len = enc.seq_len
        embedding = self.variables[_EMB]
        init = self.bridge(enc)
        helper = TrainingHelper(embedding, labels, labels_seq_len - 1)
        out_train, _ = self.decoder(init, helper, keep_prob=float(keep_prob_decoder), is_training=is_training,
                                    keep_prob_embedding=float(keep_prob_embedding), dropout_seed=self._step * 7 + 3)
        B = inputs.shape[0]

        def infer():
            saved = self.decoder._saved                  # keep the training pass's tensors
            h = GreedyEmbeddingHelper(embedding, torch.full((B,), self.sos_index, dtype=torch.int32,
                                                            device=self.device), self.eos_index)
            out, _ = self.decoder(init, h)
            self.decoder._saved = saved
            return out
        self._enc_out, self._init_state = enc, init
        return out_train.logits, out_train, _LazyDecoderOutput(infer), enc.outputs

    @graph_op(n_out=4, name="compute_loss")
    def compute_loss(self, inputs, labels, inputs_seq_len, labels_seq_len, keep_prob_encoder,
                     keep_prob_decoder, keep_prob_embedding, scope=None, is_training=True):
        """-> (total_loss, logits [B,T_out-1,V], decoder_outputs_train, decoder_outputs_infer)
        (attention_seq2seq.py:579-664)"""
        inputs = self._dev(inputs, torch.float32)
        labels = self._dev(labels, torch.int32)
        inputs_seq_len = self._dev(inputs_seq_len, torch.int32)
        labels_seq_len = self._dev(labels_seq_len, torch.int32)
        logits, out_train, out_infer, _ = self._build(
            inputs, labels, inputs_seq_len, labels_seq_len, keep_prob_encoder, keep_prob_decoder,
            keep_prob_embedding, is_training)
        total_loss, dlogits = self._sequence_loss(logits, labels, labels_seq_len, 1.0, is_training)
        total_loss = self._add_weight_decay(total_loss)
        self._ctx = {"dlogits": dlogits, "shape": tuple(inputs.shape)} if is_training else None
        return total_loss, logits, out_train, out_infer

    def _sequence_loss(self, logits, labels, labels_seq_len, scale, need_grad):
        """masked mean cross entropy over labels[:, 1:] (:619-636), gradient scaled by ``scale``"""
        L = logits.shape[1]
        if L == 0:
            return torch.zeros((), device=self.device), None
        loss, dlogits = ops.sequence_loss(logits.contiguous(), labels[:, 1:], labels_seq_len - 1,
                                          temperature=self.logits_temperature, grad_scale=scale,
                                          need_grad=need_grad)
        self.sequence_loss = loss
        return loss * scale, dlogits

    def _add_weight_decay(self, total_loss):
        if self.weight_decay > 0:
            sq = ops.clip_by_norm_multi(self._decay_params, 3.0e38)       # squared norms, no scaling
            total_loss = total_loss + 0.5 * float(self.weight_decay) * sq.sum()
        return total_loss

    def decode_beam(self, inputs, inputs_seq_len, beam_width=20, length_penalty_weight=0.6):
        """Beam-search inference (the reference's ``_beam_search_decoder_wrapper``, :550-577, is dead
        code limited to one utterance; here every utterance carries its own beam).
        -> (ids [B,W,L'], lengths [B,W], log_probs [B,W], scores [B,W]); beam 0 is the best."""
        assert isinstance(beam_width, int) and beam_width >= 1
        inputs = self._dev(inputs, torch.float32)
        inputs_seq_len = self._dev(inputs_seq_len, torch.int32)
        enc = self._encode(inputs, inputs_seq_len, 1.0, is_training=False)
        self.decoder.encoder_outputs, self.decoder.encoder_outputs_seq_len = enc.outputs, enc.seq_len
        init = self.bridge(enc)
        return self.decoder.beam_search(init, self.variables[_EMB], self.sos_index, self.eos_index, beam_width,
                                        length_penalty_weight)

    # -------------------------------------------------------------- backward
    def _backward(self):
        assert self._ctx is not None, "train() needs a preceding compute_loss(is_training=True)"
        B, T, _ = self._ctx["shape"]
        H = self.encoder_num_units
        self.flat_grads.zero_()
        d_enc = torch.zeros((B, T, 2 * H), dtype=torch.float32, device=self.device)
        dc0, dh0 = self.decoder.backward(self._ctx["dlogits"], self._dec_grads, self._cell_grads,
                                         self._att_grads, self.grads[_EMB], d_enc)
        # bridge (bridge.py:128-151): [c0, h0] = flat . W + b
        d_out = torch.cat([dc0, dh0], dim=1)
        fs = self._enc_out.final_state
        flat = torch.cat([fs[0][0], fs[0][1], fs[1][0], fs[1][1]], dim=1).contiguous()
        ops.gemm(flat, d_out, True, False, out=self._bridge_grads["bridge/weights"], beta=1.0)
        ops.colsum(d_out, out=self._bridge_grads["bridge/biases"], accumulate=True)
        d_flat = ops.gemm(d_out, self.bridge.variables["bridge/weights"], False, True)    # [B, 4H]
        d_final = ops.transpose_01(d_flat.view(B, 4, H))                                  # [4, B, H]
        d_enc_tm = ops.transpose_01(d_enc)
        d_enc_tm = self._backward_extra(d_enc_tm)
        self.encoder.backward(d_enc_tm, self._enc_vars, self._enc_grads, d_final_state=d_final)
        if self.weight_decay > 0:
            ops.axpy_multi(self._decay_params, self._decay_grads, float(self.weight_decay))
        self._ctx = None

    def _backward_extra(self, d_enc_tm):
        return d_enc_tm

    # ---------------------------------------------------------------- decode
    @graph_op(n_out=2, name="decode")
    def decode(self, decoder_outputs_train, decoder_outputs_infer):
        """-> (decoded_train [B,T_out-1], decoded_infer [B,<=max_decode_length])  (:666-699)"""
        return decoder_outputs_train.predicted_ids, decoder_outputs_infer.predicted_ids

    @graph_op(name="compute_ler")
    def compute_ler(self, labels_true, labels_pred):
        """mean_b edit_distance(pred_b, true_b) / len(true_b)  (:701-724); sparse triples or lists"""
        from ..ctc.ctc import ler_from_lists
        from ...utils.io.labels.sparsetensor import sparse_to_label_lists

        def lists(x):
            if hasattr(x, "dense_shape"):
                return sparse_to_label_lists(x, int(x.dense_shape[0]))
            if isinstance(x, (list, tuple)) and len(x) == 3 and getattr(x[0], "ndim", 0) == 2:
                return sparse_to_label_lists(x, int(np.asarray(x[2])[0]))
            return [list(r) for r in x]
        t, p = lists(labels_true), lists(labels_pred)
        return ler_from_lists(p, t, self.device)
