"""ctypes binding of the C ABI declared in ``include/b2asr.h``.

The shared library is the product's only arithmetic back-end.  There is no CPU
fallback: if ``libb2asr.so`` is missing or a call fails, a ``RuntimeError`` is
raised (see INTEGRATION.md for the build step).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2ASR_LIB", os.path.join(_HERE, "libb2asr.so"))   # override: A/B builds

PREC_FP32, PREC_BF16 = 0, 1
OPT_KINDS = {"sgd": 0, "momentum": 1, "nestrov": 2, "adagrad": 3, "adadelta": 4, "adam": 5,
             "rmsprop": 6}


class LstmDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("B", C.c_int32), ("D_in", C.c_int32), ("H", C.c_int32),
                ("use_peephole", C.c_int32), ("forget_bias", C.c_float), ("cell_clip", C.c_float),
                ("keep_prob", C.c_float), ("dropout_seed", C.c_uint64), ("precision", C.c_int32),
                ("need_backward", C.c_int32), ("num_proj", C.c_int32),
                # backward-only hand-over of the dropout mask between stacked layers (include/b2asr.h)
                ("dx_keep_prob", C.c_float), ("dx_dropout_seed", C.c_uint64), ("dy_premasked", C.c_int32)]


class LstmParams(C.Structure):
    _fields_ = [("kernel", C.c_void_p), ("bias", C.c_void_p), ("w_i_diag", C.c_void_p),
                ("w_f_diag", C.c_void_p), ("w_o_diag", C.c_void_p), ("projection", C.c_void_p)]


LstmGrads = LstmParams  # same layout (non-const pointers)


class GruDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("B", C.c_int32), ("D_in", C.c_int32), ("H", C.c_int32), ("keep_prob", C.c_float),
                ("dropout_seed", C.c_uint64), ("need_backward", C.c_int32)]


class GruParams(C.Structure):
    _fields_ = [("gates_kernel", C.c_void_p), ("gates_bias", C.c_void_p), ("cand_kernel", C.c_void_p),
                ("cand_bias", C.c_void_p)]


class VggDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("keep_prob", C.c_float),
                ("dropout_seed", C.c_uint64), ("precision", C.c_int32)]


class VggParams(C.Structure):
    _fields_ = [("conv_w", C.c_void_p * 4), ("conv_b", C.c_void_p * 4), ("fc_w", C.c_void_p),
                ("fc_b", C.c_void_p)]


VggGrads = VggParams


class DecoderDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("E", C.c_int32), ("Hd", C.c_int32), ("A", C.c_int32),
                ("emb", C.c_int32), ("C", C.c_int32), ("attention_mode", C.c_int32),
                ("query_projected", C.c_int32), ("filter_width", C.c_int32), ("sharpening", C.c_float),
                ("sigmoid_smoothing", C.c_int32), ("forget_bias", C.c_float), ("cell_clip", C.c_float),
                ("feed_previous_attention", C.c_int32), ("keep_prob_decoder", C.c_float),
                ("keep_prob_embedding", C.c_float), ("dropout_seed", C.c_uint64)]


_DEC_FIELDS = ("cell_kernel", "cell_bias", "w_i_diag", "w_f_diag", "w_o_diag", "w_query", "conv_filter",
               "w_filter", "b_filter", "v_a", "w_av", "w_out", "b_out", "embedding")


class DecoderParams(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in _DEC_FIELDS]


DecoderGrads = DecoderParams

_p, _i, _f, _sz, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64

# name -> (restype, argtypes); mirrors include/b2asr.h one to one
PROTOTYPES = {
    "b2_version": (_i, []),
    "b2_last_error": (C.c_char_p, []),
    "b2_device_is_sm100": (_i, []),
    "b2_launch_count": (C.c_ulonglong, []),
    "b2_ctc_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "b2_ctc_loss_grad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _sz, _p]),
    "b2_ctc_greedy_decode": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "b2_ctc_beam_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "b2_ctc_beam_decode": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "b2_ctc_beam_tf_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "b2_ctc_beam_decode_tf": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "b2_softmax_rows": (_i, [_p, _p, _i64, _i, _p]),
    "b2_gemm_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "b2_gemm": (_i, [_i, _i, _i, _i, _i, _f, _p, _i, _p, _i, _f, _p, _i, _p, _i, _p, _sz, _p]),
    "b2_gemm_lp": (_i, [_i, _i, _i, _i, _i, _f, _p, _i, _p, _i, _p, _i, _f, _p, _i, _p, _i, _p, _sz, _p]),
    "b2_gemm_bf16": (_i, [_i, _i, _i, _i, _i, _f, _p, _i, _p, _i, _p, _i, _p, _i, _i, _p]),
    "b2_crc32c": (C.c_uint32, [C.c_uint32, _p, _sz]),
    "b2_bgru_reserve_bytes": (_sz, [C.POINTER(GruDesc)]),
    "b2_bgru_workspace_bytes": (_sz, [C.POINTER(GruDesc)]),
    "b2_bgru_layer_forward": (_i, [C.POINTER(GruDesc), _p, _p, C.POINTER(GruParams), C.POINTER(GruParams), _p, _p, _p,
                                   _p, _sz, _p]),
    "b2_bgru_layer_backward": (_i, [C.POINTER(GruDesc), _p, _p, C.POINTER(GruParams), C.POINTER(GruParams), _p, _p, _p,
                                    C.POINTER(GruParams), C.POINTER(GruParams), _p, _sz, _p]),
    "b2_blstm_layer_path": (_i, [C.POINTER(LstmDesc)]),
    "b2_blstm_reserve_bytes": (_sz, [C.POINTER(LstmDesc)]),
    "b2_blstm_workspace_bytes": (_sz, [C.POINTER(LstmDesc)]),
    "b2_blstm_layer_forward": (_i, [C.POINTER(LstmDesc), _p, _p, _p, C.POINTER(LstmParams),
                                    C.POINTER(LstmParams), _p, _p, _p, _p, _sz, _p]),
    "b2_blstm_backward_join": (_i, [_p]),
    "b2_blstm_backward_side_wait": (_i, [_p]),
    "b2_blstm_profile_enable": (None, [_i]),
    "b2_blstm_profile_last_ms": (_i, [_p, _p]),
    "b2_blstm_reserve_y_lp": (_p, [C.POINTER(LstmDesc), _p]),
    "b2_blstm_layer_backward": (_i, [C.POINTER(LstmDesc), _p, _p, _p, C.POINTER(LstmParams),
                                     C.POINTER(LstmParams), _p, _p, _p, C.POINTER(LstmGrads),
                                     C.POINTER(LstmGrads), _p, _sz, _p]),
    "b2_blstm_layer_backward_ex": (_i, [C.POINTER(LstmDesc), _p, _p, _p, C.POINTER(LstmParams),
                                        C.POINTER(LstmParams), _p, _p, _p, _p, C.POINTER(LstmGrads),
                                        C.POINTER(LstmGrads), _p, _sz, _p]),
    "b2_attention_decoder_reserve_bytes": (_sz, [C.POINTER(DecoderDesc), _i]),
    "b2_attention_decoder_workspace_bytes": (_sz, [C.POINTER(DecoderDesc), _i]),
    "b2_attention_decoder_forward": (_i, [C.POINTER(DecoderDesc), C.POINTER(DecoderParams), _p, _p, _p, _p, _p, _p,
                                          _i, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                          C.POINTER(C.c_int32), _p, _sz, _p]),
    "b2_attention_decoder_backward": (_i, [C.POINTER(DecoderDesc), C.POINTER(DecoderParams), _p, _p, _p, _p, _i, _i,
                                           _p, _p, C.POINTER(DecoderGrads), _p, _p, _p, _p, _p, _sz, _p]),
    "b2_attention_decoder_beam_workspace_bytes": (_sz, [C.POINTER(DecoderDesc), _i, _i]),
    "b2_attention_decoder_beam_search": (_i, [C.POINTER(DecoderDesc), C.POINTER(DecoderParams), _p, _p, _p, _p, _p,
                                              _i, _i, _i, _f, _i, _i, _p, _p, _p, _p, C.POINTER(C.c_int32),
                                              _p, _sz, _p]),
    "b2_edit_distance": (_i, [_p, _p, _p, _p, _i, _i, _p, _p]),
    "b2_relu_dropout_forward": (_i, [_p, _i64, _f, C.c_uint64, _p]),
    "b2_relu_dropout_backward": (_i, [_p, _p, _i64, _f, _p, _p]),
    "b2_stack_splice_out_dim": (_i, [_i, _i, _i]),
    "b2_stack_splice": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "b2_vgg_reserve_bytes": (_sz, [C.POINTER(VggDesc)]),
    "b2_vgg_workspace_bytes": (_sz, [C.POINTER(VggDesc)]),
    "b2_vgg_output_size": (_i, [C.POINTER(VggDesc)]),
    "b2_vgg_frontend_forward": (_i, [C.POINTER(VggDesc), _p, C.POINTER(VggParams), _p, _p, _p, _sz, _p]),
    "b2_vgg_frontend_backward": (_i, [C.POINTER(VggDesc), C.POINTER(VggParams), _p, _p, C.POINTER(VggGrads),
                                      _p, _sz, _p]),
    "b2_attention_step_forward": (_i, [_i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _f,
                                       _i, _p, _p, _p, _p]),
    "b2_attention_step_backward_workspace_bytes": (_sz, [_i, _i]),
    "b2_attention_step_backward": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i,
                                        _p, _p, _p, _i, _p, _p, _p, _sz, _p]),
    "b2_attention_step_backward_loc": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i,
                                            _p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "b2_sequence_loss": (_i, [_p, _p, _i, _p, _i, _i, _i, _f, _f, _p, _p, _p]),
    "b2_tanh_backward": (_i, [_p, _p, _p, _i64, _p]),
    "b2_lstm_cell_pointwise_backward": (_i, [_p] * 8 + [_i, _i, _f, _f, _p, _p, _p]),
    "b2_decoder_peephole_grad": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "b2_dropout_rows": (_i, [_p, _i, _p, _i, _i64, _i, _f, C.c_uint64, C.c_uint64, C.c_uint64, _i, _p]),
    "b2_embedding_grad": (_i, [_p, _i, _p, _i64, _i, _i, _p, _p]),
    "b2_lstm_cell_pointwise": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _f, _f, _p, _p, _p]),
    "b2_tanh_inplace": (_i, [_p, _i64, _p]),
    "b2_decoder_step_emit": (_i, [_i] * 8 + [_p] * 12 + [_i, _p, _i, _i] + [_p] * 6 + [_p]),
    "b2_argmax_rows": (_i, [_p, _i64, _i, _p, _p]),
    "b2_transpose_01": (_i, [_p, _p, _i, _i, _i, _p]),
    "b2_colsum": (_i, [_p, _i64, _i, _i, _p, _i, _p]),
    "b2_clip_by_norm_multi": (_i, [_p, _p, _i, _f, _f, _p, _p]),
    "b2_axpy_multi": (_i, [_p, _p, _p, _i, _f, _p]),
    "b2_tower_mean": (_i, [_p, _i, _p, _i64, _p]),
    "b2_optimizer_step_multi": (_i, [_i, _p, _p, _p, _p, _p, _i, _f, _i64, _p]),
    "b2_comm_available": (_i, []),
    "b2_comm_get_unique_id": (_i, [_p]),
    "b2_comm_init_rank": (_i, [C.POINTER(C.c_void_p), _i, _p, _i]),
    "b2_comm_init_all": (_i, [C.POINTER(C.c_void_p), _i, C.POINTER(C.c_int)]),
    "b2_comm_size": (_i, [_p]),
    "b2_comm_destroy": (_i, [_p]),
    "b2_allreduce_mean": (_i, [_p, _p, _p, _i, _p]),
    "b2_allreduce_mean_local": (_i, [_p, _p, _i64, _i, _p]),
}

_lib = None


def load():
    """Load libb2asr.so (once) and attach prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libb2asr.so not found at %s -- build it with "
            "`python -m tensorflow_end2end_speech_recognition_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().b2_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))
