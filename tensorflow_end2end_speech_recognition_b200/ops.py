"""Tensor-level wrappers over the C ABI.  torch tensors are only containers
(device memory + stream); every FLOP happens in libb2asr.so."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import LstmDesc, LstmParams, PREC_BF16, PREC_FP32  # noqa: F401

_ws_cache = {}


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b2asr ops need CUDA tensors (no CPU fallback)")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def workspace(name, nbytes, device):
    """Grow-only cached scratch buffer per (name, device)."""
    key = (name, str(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def pack_labels(labels):
    """list of int sequences -> (flat int32, offsets int32[B+1], max_len)."""
    lens = [len(l) for l in labels]
    offs = np.zeros(len(labels) + 1, dtype=np.int32)
    offs[1:] = np.cumsum(lens)
    flat = np.zeros(max(int(offs[-1]), 1), dtype=np.int32)
    if offs[-1] > 0:
        flat[:offs[-1]] = np.concatenate([np.asarray(l, dtype=np.int32) for l in labels if len(l)])
    return flat, offs, (max(lens) if lens else 0)


def check_labels(labels, num_classes, blank=None, what="labels"):
    """tf.nn.ctc_loss rejects label values outside [0, num_classes) or equal to the blank with
    InvalidArgument (its kernel indexes the softmax row by label); same contract here, raised on the host
    before anything is launched.  A dense -1-padded row (utils/dataset/ctc.py pads with -1) must be
    stripped by the caller -- ``list2sparsetensor(..., padded_value=-1)`` does."""
    if blank is None:
        blank = num_classes - 1
    for b, l in enumerate(labels):
        if len(l) == 0:
            continue
        a = np.asarray(l)
        if a.min() < 0 or a.max() >= num_classes or (a == blank).any():
            raise ValueError("%s[%d]: label values must be in [0, %d) and differ from the blank index %d, got "
                             "min %d max %d" % (what, b, num_classes, blank, int(a.min()), int(a.max())))


def ctc_loss_grad(logits, labels_flat, label_offsets, seq_len, max_label_len, blank=None,
                  ignore_longer=True, grad_scale=1.0, need_grad=True):
    """logits [T,B,C] f32 cuda; labels_flat/label_offsets/seq_len int32 cuda.
    Returns (loss [B], grad [T,B,C] or None)."""
    lib = _lib.load()
    _require_cuda(logits, labels_flat, label_offsets, seq_len)
    assert logits.dtype == torch.float32 and logits.is_contiguous()
    T, B, Cc = logits.shape
    if blank is None:
        blank = Cc - 1
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits) if need_grad else None
    nbytes = lib.b2_ctc_workspace_bytes(T, B, Cc, int(max_label_len))
    ws = workspace("ctc", nbytes, logits.device)
    rc = lib.b2_ctc_loss_grad(_ptr(logits), _ptr(labels_flat), _ptr(label_offsets), _ptr(seq_len),
                              T, B, Cc, int(blank), int(max_label_len), int(bool(ignore_longer)),
                              float(grad_scale), _ptr(loss), _ptr(grad), _ptr(ws), nbytes, _stream())
    _lib.check(rc, "b2_ctc_loss_grad")
    return loss, grad


def ctc_greedy_decode(logits, seq_len, blank=None):
    """logits [T,B,C] -> (labels [B,T] int32 padded -1, lengths [B] int32)."""
    lib = _lib.load()
    _require_cuda(logits, seq_len)
    T, B, Cc = logits.shape
    if blank is None:
        blank = Cc - 1
    out = torch.empty((B, T), dtype=torch.int32, device=logits.device)
    n = torch.empty(B, dtype=torch.int32, device=logits.device)
    rc = lib.b2_ctc_greedy_decode(_ptr(logits.contiguous()), _ptr(seq_len), T, B, Cc, int(blank),
                                  _ptr(out), _ptr(n), _stream())
    _lib.check(rc, "b2_ctc_greedy_decode")
    return out, n


def ctc_beam_decode(log_probs, seq_len, beam_width, blank=None):
    """log_probs [B,T,C] natural-log posteriors -> (labels [B,T], lengths [B], score [B])."""
    lib = _lib.load()
    _require_cuda(log_probs, seq_len)
    B, T, Cc = log_probs.shape
    if blank is None:
        blank = Cc - 1
    out = torch.empty((B, T), dtype=torch.int32, device=log_probs.device)
    n = torch.empty(B, dtype=torch.int32, device=log_probs.device)
    score = torch.empty(B, dtype=torch.float32, device=log_probs.device)
    nbytes = lib.b2_ctc_beam_workspace_bytes(T, B, Cc, int(beam_width))
    ws = workspace("beam", nbytes, log_probs.device)
    rc = lib.b2_ctc_beam_decode(_ptr(log_probs.contiguous()), _ptr(seq_len), T, B, Cc, int(blank),
                                int(beam_width), _ptr(out), _ptr(n), _ptr(score), _ptr(ws), nbytes,
                                _stream())
    _lib.check(rc, "b2_ctc_beam_decode")
    return out, n, score


def ctc_beam_decode_tf(logits, seq_len, beam_width, blank=None, merge_repeated=True):
    """tf.nn.ctc_beam_search_decoder semantics.  logits [T,B,C] raw scores -> (labels [B,T] -1 padded, lengths [B],
    score [B])."""
    lib = _lib.load()
    _require_cuda(logits, seq_len)
    T, B, Cc = logits.shape
    if blank is None:
        blank = Cc - 1
    out = torch.empty((B, T), dtype=torch.int32, device=logits.device)
    n = torch.empty(B, dtype=torch.int32, device=logits.device)
    score = torch.empty(B, dtype=torch.float32, device=logits.device)
    nbytes = lib.b2_ctc_beam_tf_workspace_bytes(T, B, Cc, int(beam_width))
    ws = workspace("beam_tf", nbytes, logits.device)
    rc = lib.b2_ctc_beam_decode_tf(_ptr(logits.contiguous()), _ptr(seq_len), T, B, Cc, int(blank), int(beam_width),
                                   int(bool(merge_repeated)), _ptr(out), _ptr(n), _ptr(score), _ptr(ws), nbytes,
                                   _stream())
    _lib.check(rc, "b2_ctc_beam_decode_tf")
    return out, n, score


def softmax_rows(x):
    lib = _lib.load()
    _require_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    Cc = x.shape[-1]
    rc = lib.b2_softmax_rows(_ptr(x), _ptr(y), x.numel() // Cc, Cc, _stream())
    _lib.check(rc, "b2_softmax_rows")
    return y


def gemm(A, B, transa=False, transb=False, bias=None, precision=PREC_FP32, out=None, beta=0.0,
         alpha=1.0, a_lp=None):
    """C = alpha*op(A).op(B) + beta*C + bias.  2-D f32 cuda tensors (row-major, may be strided rows).
    a_lp = (device pointer, row stride) of a bf16 shadow of A the caller already holds (bf16 path only)."""
    lib = _lib.load()
    _require_cuda(A, B, bias, out)
    assert A.dtype == torch.float32 and B.dtype == torch.float32
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, K = (A.shape[1], A.shape[0]) if transa else (A.shape[0], A.shape[1])
    N = B.shape[0] if transb else B.shape[1]
    Kb = B.shape[1] if transb else B.shape[0]
    assert K == Kb, "inner dimensions differ"
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
        assert beta == 0.0
    nbytes = lib.b2_gemm_workspace_bytes(M, N, K, precision)
    ws = workspace("gemm", nbytes, A.device) if nbytes else None
    lp_ptr, lp_ld = (a_lp if a_lp else (0, 0))
    rc = lib.b2_gemm_lp(int(transa), int(transb), M, N, K, float(alpha), _ptr(A), A.stride(0),
                        C.c_void_p(lp_ptr or 0), int(lp_ld), _ptr(B),
                        B.stride(0), float(beta), _ptr(out), out.stride(0), _ptr(bias), int(precision),
                        _ptr(ws), nbytes, _stream())
    _lib.check(rc, "b2_gemm")
    return out


def lstm_cell_pointwise(z, bias, peep, c_prev, forget_bias=1.0, cell_clip=None, out_c=None, out_h=None):
    """z [B,4H] pre-activations -> (c [B,H], h [B,H]); peep = (w_i, w_f, w_o) or None."""
    lib = _lib.load()
    _require_cuda(z, c_prev)
    B, H4 = z.shape
    H = H4 // 4
    c = out_c if out_c is not None else torch.empty((B, H), dtype=torch.float32, device=z.device)
    h = out_h if out_h is not None else torch.empty((B, H), dtype=torch.float32, device=z.device)
    wi, wf, wo = peep if peep is not None else (None, None, None)
    rc = lib.b2_lstm_cell_pointwise(_ptr(z.contiguous()), _ptr(bias), _ptr(wi), _ptr(wf), _ptr(wo),
                                    _ptr(c_prev.contiguous()), B, H, float(forget_bias),
                                    float(cell_clip) if cell_clip else 0.0, _ptr(c), _ptr(h), _stream())
    _lib.check(rc, "b2_lstm_cell_pointwise")
    return c, h


def lstm_cell_pointwise_backward(z, bias, peep, c_prev, dh, dc_in, forget_bias=1.0, cell_clip=None,
                                 out_dz=None):
    """-> (dz [B,4H], dc_prev [B,H])"""
    lib = _lib.load()
    _require_cuda(z, c_prev, dh)
    B, H4 = z.shape
    H = H4 // 4
    dz = out_dz if out_dz is not None else torch.empty((B, H4), dtype=torch.float32, device=z.device)
    dc_prev = torch.empty((B, H), dtype=torch.float32, device=z.device)
    wi, wf, wo = peep if peep is not None else (None, None, None)
    rc = lib.b2_lstm_cell_pointwise_backward(_ptr(z), _ptr(bias), _ptr(wi), _ptr(wf), _ptr(wo), _ptr(c_prev),
                                             _ptr(dh), _ptr(dc_in), B, H, float(forget_bias),
                                             float(cell_clip) if cell_clip else 0.0, _ptr(dz), _ptr(dc_prev),
                                             _stream())
    _lib.check(rc, "b2_lstm_cell_pointwise_backward")
    return dz, dc_prev


def sequence_loss(logits, targets, lengths, temperature=1.0, grad_scale=1.0, need_grad=True):
    """logits [B,L,C] f32, targets [B,>=L] int32 view (row stride = its stride(0)), lengths [B]
    -> (loss 0-d tensor, dlogits [B,L,C] or None)."""
    lib = _lib.load()
    _require_cuda(logits, targets, lengths)
    B, L, Cc = logits.shape
    assert logits.is_contiguous() and targets.stride(1) == 1 and targets.dtype == torch.int32
    rowloss = torch.empty((B, L), dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits) if need_grad else None
    rc = lib.b2_sequence_loss(_ptr(logits), _ptr(targets), targets.stride(0), _ptr(lengths), B, L, Cc,
                              float(temperature), float(grad_scale), _ptr(rowloss), _ptr(dlogits), _stream())
    _lib.check(rc, "b2_sequence_loss")
    wsum = torch.clamp(lengths, 0, L).sum().to(torch.float32)
    return rowloss.sum() / (wsum + 1e-12), dlogits


def tanh_backward(dy, y, out=None):
    lib = _lib.load()
    _require_cuda(dy, y)
    assert dy.is_contiguous() and y.is_contiguous()
    dx = out if out is not None else torch.empty_like(dy)
    _lib.check(lib.b2_tanh_backward(_ptr(dy), _ptr(y), _ptr(dx), dy.numel(), _stream()), "b2_tanh_backward")
    return dx


def decoder_peephole_grad(dz_all, c_all, steps, B, H, dwi, dwf, dwo):
    lib = _lib.load()
    rc = lib.b2_decoder_peephole_grad(_ptr(dz_all), _ptr(c_all), steps, B, H, _ptr(dwi), _ptr(dwf), _ptr(dwo),
                                      _stream())
    _lib.check(rc, "b2_decoder_peephole_grad")


def embedding_grad(dx, ldx, ids, rows, D, dW):
    lib = _lib.load()
    rc = lib.b2_embedding_grad(_ptr(dx), ldx, _ptr(ids), rows, D, dW.shape[0], _ptr(dW), _stream())
    _lib.check(rc, "b2_embedding_grad")


def tanh_(x):
    lib = _lib.load()
    _require_cuda(x)
    assert x.is_contiguous()
    _lib.check(lib.b2_tanh_inplace(_ptr(x), x.numel(), _stream()), "b2_tanh_inplace")
    return x


def argmax_rows(x):
    lib = _lib.load()
    _require_cuda(x)
    x = x.contiguous()
    rows, Cc = x.shape
    out = torch.empty(rows, dtype=torch.int32, device=x.device)
    _lib.check(lib.b2_argmax_rows(_ptr(x), rows, Cc, _ptr(out), _stream()), "b2_argmax_rows")
    return out


def transpose_01(x):
    lib = _lib.load()
    _require_cuda(x)
    x = x.contiguous()
    d0, d1, d2 = x.shape
    y = torch.empty((d1, d0, d2), dtype=x.dtype, device=x.device)
    _lib.check(lib.b2_transpose_01(_ptr(x), _ptr(y), d0, d1, d2, _stream()), "b2_transpose_01")
    return y


def colsum(X, out=None, accumulate=False):
    lib = _lib.load()
    _require_cuda(X)
    M, N = X.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=X.device)
        accumulate = False
    _lib.check(lib.b2_colsum(_ptr(X), M, N, X.stride(0), _ptr(out), int(accumulate), _stream()),
               "b2_colsum")
    return out


# --------------------------------------------------------------------------- LSTM
def _params_struct(p):
    return LstmParams(p["kernel"].data_ptr(), p["bias"].data_ptr(),
                      p["w_i_diag"].data_ptr() if "w_i_diag" in p else 0,
                      p["w_f_diag"].data_ptr() if "w_f_diag" in p else 0,
                      p["w_o_diag"].data_ptr() if "w_o_diag" in p else 0,
                      p["projection"].data_ptr() if "projection" in p else 0)


def lstm_desc(T, B, D_in, H, use_peephole=True, forget_bias=1.0, cell_clip=None, keep_prob=1.0,
              dropout_seed=0, precision=PREC_FP32, need_backward=True, num_proj=None):
    return LstmDesc(T, B, D_in, H, int(bool(use_peephole)), float(forget_bias),
                    float(cell_clip) if cell_clip else 0.0, float(keep_prob), int(dropout_seed),
                    int(precision), int(bool(need_backward)), int(num_proj or 0), 0.0, 0, 0)


def lstm_desc_with(desc, **fields):
    """copy of a layer descriptor with some fields replaced (dx_keep_prob, dx_dropout_seed, dy_premasked, ...)"""
    d = type(desc).from_buffer_copy(desc)
    for k, v in fields.items():
        setattr(d, k, v)
    return d


def blstm_layer_path(desc):
    """0: fp32 / hybrid step kernels, 1: cluster/TMEM tcgen05 recurrence, 2: grid-resident wide-layer recurrence"""
    return int(_lib.load().b2_blstm_layer_path(C.byref(desc)))


def blstm_layer_forward(desc, x, seq_len, p_fw, p_bw, want_final_state=False, x_lp=0):
    """x [T,B,D] -> (y [T,B,2H], final_state [4,B,H] or None, reserve buffer).
    x_lp: raw device pointer of a bf16 shadow of x (from reserve_y_lp of the layer below) or 0.
    With desc.num_proj = P: y [T,B,2P]; final_state is the tuple (c_fw [B,H], h_fw [B,P], c_bw, h_bw)."""
    lib = _lib.load()
    _require_cuda(x, seq_len)
    dev = x.device
    Hout = desc.num_proj if desc.num_proj > 0 else desc.H
    y = torch.empty((desc.T, desc.B, 2 * Hout), dtype=torch.float32, device=dev)
    fs = None
    if want_final_state:
        fs = torch.empty((4, desc.B, desc.H), dtype=torch.float32, device=dev) if desc.num_proj <= 0 else \
            torch.empty(2 * desc.B * (desc.H + Hout), dtype=torch.float32, device=dev)
    # the bf16 path keeps its bf16 output shadow in `reserve` even for inference
    reserve = torch.empty(lib.b2_blstm_reserve_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
    nbytes = lib.b2_blstm_workspace_bytes(C.byref(desc))
    ws = workspace("lstm", nbytes, dev)
    fw, bw = _params_struct(p_fw), _params_struct(p_bw)
    rc = lib.b2_blstm_layer_forward(C.byref(desc), _ptr(x.contiguous()), C.c_void_p(x_lp or 0),
                                    _ptr(seq_len), C.byref(fw),
                                    C.byref(bw), _ptr(y), _ptr(fs), _ptr(reserve), _ptr(ws), nbytes,
                                    _stream())
    _lib.check(rc, "b2_blstm_layer_forward")
    if fs is not None and desc.num_proj > 0:
        B, H, P = desc.B, desc.H, Hout
        o = [0, B * H, B * H + B * P, 2 * B * H + B * P]
        fs = (fs[o[0]:o[1]].view(B, H), fs[o[1]:o[2]].view(B, P), fs[o[2]:o[3]].view(B, H), fs[o[3]:].view(B, P))
    return y, fs, reserve


# ---------------------------------------------------------------- GRU layers (csrc/gru.cu)
def gru_desc(T, B, D_in, H, keep_prob=1.0, dropout_seed=0, need_backward=True):
    return _lib.GruDesc(T, B, D_in, H, float(keep_prob), int(dropout_seed), int(bool(need_backward)))


def _gru_struct(p):
    return _lib.GruParams(p["gates/kernel"].data_ptr(), p["gates/bias"].data_ptr(), p["candidate/kernel"].data_ptr(),
                          p["candidate/bias"].data_ptr())


def bgru_layer_forward(desc, x, seq_len, p_fw, p_bw, want_final_state=False):
    """x [T,B,D] -> (y [T,B,2H], final_state [2,B,H] or None, reserve).  Parameter dicts: TF GRUCell variable names
    ``gates/kernel [(D+H),2H]``, ``gates/bias``, ``candidate/kernel [(D+H),H]``, ``candidate/bias``."""
    lib = _lib.load()
    _require_cuda(x, seq_len)
    dev = x.device
    y = torch.empty((desc.T, desc.B, 2 * desc.H), dtype=torch.float32, device=dev)
    fs = torch.empty((2, desc.B, desc.H), dtype=torch.float32, device=dev) if want_final_state else None
    reserve = torch.empty(lib.b2_bgru_reserve_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
    nbytes = lib.b2_bgru_workspace_bytes(C.byref(desc))
    ws = workspace("gru", nbytes, dev)
    fw, bw = _gru_struct(p_fw), _gru_struct(p_bw)
    rc = lib.b2_bgru_layer_forward(C.byref(desc), _ptr(x.contiguous()), _ptr(seq_len), C.byref(fw), C.byref(bw),
                                   _ptr(y), _ptr(fs), _ptr(reserve), _ptr(ws), nbytes, _stream())
    _lib.check(rc, "b2_bgru_layer_forward")
    return y, fs, reserve


def bgru_layer_backward(desc, x, seq_len, p_fw, p_bw, dy, reserve, g_fw, g_bw, need_dx=True):
    """Accumulates into the gradient dicts; returns dx [T,B,D] or None."""
    lib = _lib.load()
    _require_cuda(x, dy)
    dev = x.device
    dx = torch.empty((desc.T, desc.B, desc.D_in), dtype=torch.float32, device=dev) if need_dx else None
    nbytes = lib.b2_bgru_workspace_bytes(C.byref(desc))
    ws = workspace("gru", nbytes, dev)
    fw, bw, gf, gb = _gru_struct(p_fw), _gru_struct(p_bw), _gru_struct(g_fw), _gru_struct(g_bw)
    rc = lib.b2_bgru_layer_backward(C.byref(desc), _ptr(x.contiguous()), _ptr(seq_len), C.byref(fw), C.byref(bw),
                                    _ptr(dy.contiguous()), _ptr(reserve), _ptr(dx), C.byref(gf), C.byref(gb), _ptr(ws),
                                    nbytes, _stream())
    _lib.check(rc, "b2_bgru_layer_backward")
    return dx


def blstm_backward_join():
    """current stream waits for the side-stream weight-gradient GEMMs of the bf16 path."""
    _lib.check(_lib.load().b2_blstm_backward_join(_stream()), "b2_blstm_backward_join")


def blstm_backward_side_wait(stream=None):
    """`stream` (default: current) waits for the side-stream weight-gradient GEMMs enqueued so far."""
    s = C.c_void_p(stream.cuda_stream) if stream is not None else _stream()
    _lib.check(_lib.load().b2_blstm_backward_side_wait(s), "b2_blstm_backward_side_wait")


def reserve_y_lp(desc, reserve):
    """raw device pointer (int) of the bf16 layer output kept in `reserve`, or 0."""
    if reserve is None:
        return 0
    return _lib.load().b2_blstm_reserve_y_lp(C.byref(desc), _ptr(reserve)) or 0


def blstm_layer_backward(desc, x, seq_len, p_fw, p_bw, dy, reserve, g_fw, g_bw, need_dx=True, x_lp=0,
                         d_final_state=None):
    """Accumulates into the gradient dicts g_fw / g_bw; returns dx [T,B,D] or None.
    d_final_state: [4,B,H] gradient of (c_fw, h_fw, c_bw, h_bw) or None."""
    lib = _lib.load()
    _require_cuda(x, dy)
    dev = x.device
    dx = torch.empty((desc.T, desc.B, desc.D_in), dtype=torch.float32, device=dev) if need_dx else None
    nbytes = lib.b2_blstm_workspace_bytes(C.byref(desc))
    ws = workspace("lstm", nbytes, dev)
    fw, bw = _params_struct(p_fw), _params_struct(p_bw)
    gf, gb = _params_struct(g_fw), _params_struct(g_bw)
    if d_final_state is not None:
        _require_cuda(d_final_state)
        assert d_final_state.is_contiguous() and tuple(d_final_state.shape) == (4, desc.B, desc.H)
    rc = lib.b2_blstm_layer_backward_ex(C.byref(desc), _ptr(x.contiguous()), C.c_void_p(x_lp or 0),
                                        _ptr(seq_len), C.byref(fw),
                                        C.byref(bw), _ptr(dy.contiguous()), _ptr(d_final_state),
                                        _ptr(reserve), _ptr(dx),
                                        C.byref(gf), C.byref(gb), _ptr(ws), nbytes, _stream())
    _lib.check(rc, "b2_blstm_layer_backward")
    return dx


def edit_distance(hyp_lists, ref_lists, device):
    """Levenshtein distances of paired label lists -> int32 numpy [B] (computed on the device)."""
    lib = _lib.load()
    B = len(hyp_lists)
    hf, ho, _ = pack_labels(hyp_lists)
    rf, ro, rmax = pack_labels(ref_lists)
    dev = torch.device(device)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    hf_d, ho_d, rf_d, ro_d = t(hf if len(hf) else [0]), t(ho), t(rf if len(rf) else [0]), t(ro)
    dist = torch.empty(B, dtype=torch.int32, device=dev)
    rc = lib.b2_edit_distance(_ptr(hf_d), _ptr(ho_d), _ptr(rf_d), _ptr(ro_d), B, int(rmax), _ptr(dist), _stream())
    _lib.check(rc, "b2_edit_distance")
    return dist.cpu().numpy()


def relu_dropout_(x, keep_prob=1.0, seed=0):
    lib = _lib.load()
    _require_cuda(x)
    assert x.is_contiguous()
    _lib.check(lib.b2_relu_dropout_forward(_ptr(x), x.numel(), float(keep_prob), int(seed), _stream()),
               "b2_relu_dropout_forward")
    return x


def relu_dropout_backward(d_out, out, keep_prob=1.0):
    lib = _lib.load()
    _require_cuda(d_out, out)
    d_in = torch.empty_like(out)
    _lib.check(lib.b2_relu_dropout_backward(_ptr(d_out.contiguous()), _ptr(out), out.numel(), float(keep_prob),
                                            _ptr(d_in), _stream()), "b2_relu_dropout_backward")
    return d_in


# ------------------------------------------------------------------ VGG front-end
VGG_CONVS = ("VGG1/conv1", "VGG1/conv2", "VGG2/conv1", "VGG2/conv2")


def vgg_desc(N, H, W, keep_prob=1.0, dropout_seed=0, precision=PREC_FP32):
    return _lib.VggDesc(int(N), int(H), int(W), float(keep_prob), int(dropout_seed), int(precision))


def _vgg_struct(p):
    s = _lib.VggParams()
    for i, c in enumerate(VGG_CONVS):
        s.conv_w[i] = p[c + "/weight"].data_ptr()
        s.conv_b[i] = p[c + "/bias"].data_ptr()
    s.fc_w = p["bridge/weights"].data_ptr()
    s.fc_b = p["bridge/biases"].data_ptr()
    return s


def vgg_frontend_forward(desc, x, params):
    """x [N,H,W,3] (any view of that many contiguous floats) -> (out [N,256], reserve)"""
    lib = _lib.load()
    _require_cuda(x)
    dev = x.device
    out = torch.empty((desc.N, 256), dtype=torch.float32, device=dev)
    reserve = torch.empty(lib.b2_vgg_reserve_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
    nbytes = lib.b2_vgg_workspace_bytes(C.byref(desc))
    ws = workspace("vgg", nbytes, dev)
    ps = _vgg_struct(params)
    rc = lib.b2_vgg_frontend_forward(C.byref(desc), _ptr(x.contiguous()), C.byref(ps), _ptr(out), _ptr(reserve),
                                     _ptr(ws), nbytes, _stream())
    _lib.check(rc, "b2_vgg_frontend_forward")
    return out, reserve


def vgg_frontend_backward(desc, params, d_out, reserve, grads):
    lib = _lib.load()
    _require_cuda(d_out)
    nbytes = lib.b2_vgg_workspace_bytes(C.byref(desc))
    ws = workspace("vgg", nbytes, d_out.device)
    ps, gs = _vgg_struct(params), _vgg_struct(grads)
    rc = lib.b2_vgg_frontend_backward(C.byref(desc), C.byref(ps), _ptr(d_out.contiguous()), _ptr(reserve),
                                      C.byref(gs), _ptr(ws), nbytes, _stream())
    _lib.check(rc, "b2_vgg_frontend_backward")


# ------------------------------------------------------------------ clip + optimizer
class TensorList(object):
    """Device-side arrays of pointers / sizes for the multi-tensor kernels."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        dev = self.tensors[0].device
        self.n = len(self.tensors)
        self.ptrs = torch.tensor([t.data_ptr() for t in self.tensors], dtype=torch.int64, device=dev)
        self.sizes = torch.tensor([t.numel() for t in self.tensors], dtype=torch.int64, device=dev)


def clip_by_norm_multi(grads, clip_norm, post_scale=1.0):
    """grads: TensorList.  In-place per-tensor tf.clip_by_norm, then *post_scale."""
    lib = _lib.load()
    norms = torch.empty(grads.n, dtype=torch.float32, device=grads.ptrs.device)
    rc = lib.b2_clip_by_norm_multi(_ptr(grads.ptrs), _ptr(grads.sizes), grads.n,
                                   float(clip_norm) if clip_norm else 0.0, float(post_scale),
                                   _ptr(norms), _stream())
    _lib.check(rc, "b2_clip_by_norm_multi")
    return norms


def axpy_multi(xs, ys, alpha):
    """ys[k] += alpha * xs[k]  (TensorLists with equal sizes)."""
    lib = _lib.load()
    _lib.check(lib.b2_axpy_multi(_ptr(xs.ptrs), _ptr(ys.ptrs), _ptr(xs.sizes), xs.n, float(alpha), _stream()),
               "b2_axpy_multi")


def add_(y, x, alpha=1.0):
    """y += alpha * x  (same-shape contiguous fp32 cuda tensors), in place; returns y"""
    _require_cuda(y, x)
    assert y.is_contiguous() and x.is_contiguous() and y.numel() == x.numel()
    axpy_multi(TensorList([x]), TensorList([y]), alpha)
    return y


def tower_mean(srcs, dst):
    """dst = mean of the equally-shaped fp32 cuda tensors in ``srcs`` (dst may be srcs[0])."""
    lib = _lib.load()
    _require_cuda(dst, *srcs)
    n = dst.numel()
    assert all(t.numel() == n and t.dtype == torch.float32 and t.is_contiguous() for t in srcs)
    arr = (C.c_void_p * len(srcs))(*[t.data_ptr() for t in srcs])
    _lib.check(lib.b2_tower_mean(arr, len(srcs), _ptr(dst), n, _stream()), "b2_tower_mean")
    return dst


def optimizer_step_multi(kind, params, grads, state0, state1, learning_rate, step):
    lib = _lib.load()
    rc = lib.b2_optimizer_step_multi(_lib.OPT_KINDS[kind], _ptr(params.ptrs), _ptr(grads.ptrs),
                                     _ptr(state0.ptrs) if state0 is not None else C.c_void_p(0),
                                     _ptr(state1.ptrs) if state1 is not None else C.c_void_p(0),
                                     _ptr(params.sizes), params.n, float(learning_rate), int(step),
                                     _stream())
    _lib.check(rc, "b2_optimizer_step_multi")
