// Small kernels of the attention decoder step (forward):
//   b2_lstm_cell_pointwise  gate math of one LSTMBlockCell step on pre-activations z = [x,h].W
//                           (reference: attention_seq2seq.py:352-363, cell equations
//                            models/recurrent/layers/lstm.py:142-183)
//   b2_tanh_inplace         attentional vector tanh(W[out; ctx])      (attention_decoder.py:189-196)
//   b2_decoder_step_emit    everything after the arg-max of one dynamic_decode iteration, fused:
//                           impute_finished (zero outputs / copy state through), write of the
//                           step's outputs into batch-major slot t, helper.next_inputs
//                           (embedding gather + input feeding) and the finished flags
//                                       (dynamic_decoder.py:148-196, attention_decoder.py:221-238)
//   b2_argmax_rows          GreedyEmbeddingHelper.sample = argmax(logits), first max on ties
//                                                                     (attention_seq2seq.py:490-494)
#include "common.cuh"

namespace b2 {

__global__ void __launch_bounds__(256)
lstm_cell_pointwise_kernel(const float* __restrict__ z, const float* __restrict__ bias,
                           const float* __restrict__ wi, const float* __restrict__ wf,
                           const float* __restrict__ wo, const float* __restrict__ c_prev, int B,
                           int H, float forget_bias, float cell_clip, float* __restrict__ c_out,
                           float* __restrict__ h_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, u = idx % H;
  const float* zr = z + (size_t)b * 4 * H;
  const float cp = c_prev[idx];
  float zi = zr[u] + (bias ? bias[u] : 0.f), zg = zr[H + u] + (bias ? bias[H + u] : 0.f);
  float zf = zr[2 * H + u] + (bias ? bias[2 * H + u] : 0.f) + forget_bias;
  float zo = zr[3 * H + u] + (bias ? bias[3 * H + u] : 0.f);
  if (wi) { zi += wi[u] * cp; zf += wf[u] * cp; }
  float c = sigmoidf_(zf) * cp + sigmoidf_(zi) * tanhf_(zg);
  if (cell_clip > 0.f) c = fminf(fmaxf(c, -cell_clip), cell_clip);
  if (wo) zo += wo[u] * c;
  c_out[idx] = c;
  h_out[idx] = sigmoidf_(zo) * tanhf_(c);
}

__global__ void __launch_bounds__(256) tanh_kernel(float* __restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] = tanhf_(x[i]);
}

__global__ void __launch_bounds__(256)
argmax_rows_kernel(const float* __restrict__ x, int64_t rows, int C, int* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float* xr = x + row * C;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = xr[c];
    if (v > best) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) out[row] = (bi == 0x7fffffff) ? 0 : bi;
}

struct EmitArgs {
  int B, C, Hd, E, T, emb, t, L;
  const float* logits; const int* ids; const float* av; const float* alpha; const float* ctx;
  const float* c_new; const float* h_new;
  float* c_state; float* h_state; int* finished;
  const float* embedding; const int* labels; int labels_ld; const int* dec_len; int eos; int max_iter;
  float* xh;
  float* out_logits; int* out_ids; float* out_av; float* out_alpha; float* out_ctx;
};

// one CTA per utterance
__global__ void __launch_bounds__(256) decoder_step_emit_kernel(const EmitArgs a) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int fin = a.finished[b];
  const int id = a.ids[b];
  const size_t slot = (size_t)b * a.L + a.t;
  // outputs, zeroed past finish
  for (int i = tid; i < a.C; i += 256) a.out_logits[slot * a.C + i] = fin ? 0.f : a.logits[(size_t)b * a.C + i];
  for (int i = tid; i < a.Hd; i += 256) a.out_av[slot * a.Hd + i] = fin ? 0.f : a.av[(size_t)b * a.Hd + i];
  for (int i = tid; i < a.T; i += 256) a.out_alpha[slot * a.T + i] = fin ? 0.f : a.alpha[(size_t)b * a.T + i];
  for (int i = tid; i < a.E; i += 256) a.out_ctx[slot * a.E + i] = fin ? 0.f : a.ctx[(size_t)b * a.E + i];
  if (tid == 0) a.out_ids[slot] = fin ? 0 : id;
  // state copy-through, and the cell input of the next step: [emb(next) ; ctx ; h]
  const int ldx = a.emb + a.E + a.Hd;
  float* xr = a.xh + (size_t)b * ldx;
  for (int i = tid; i < a.Hd; i += 256) {
    const size_t k = (size_t)b * a.Hd + i;
    const float c = fin ? a.c_state[k] : a.c_new[k];
    const float h = fin ? a.h_state[k] : a.h_new[k];
    a.c_state[k] = c; a.h_state[k] = h;
    xr[a.emb + a.E + i] = h;
  }
  for (int i = tid; i < a.E; i += 256) xr[a.emb + i] = a.ctx[(size_t)b * a.E + i];
  int step_fin, next_id;
  if (a.labels) {                                  // TrainingHelper
    step_fin = (a.t + 1) >= a.dec_len[b];
    next_id = (a.t + 1 < a.labels_ld - 1) ? a.labels[(size_t)b * a.labels_ld + a.t + 1] : -1;
  } else {                                         // GreedyEmbeddingHelper
    step_fin = (id == a.eos);
    next_id = id;
  }
  for (int i = tid; i < a.emb; i += 256)
    xr[i] = (next_id >= 0 && next_id < a.C) ? a.embedding[(size_t)next_id * a.emb + i] : 0.f;
  if (tid == 0) {
    int f = fin | step_fin;
    if (a.max_iter > 0 && a.t + 1 >= a.max_iter) f = 1;
    a.finished[b] = f;
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_lstm_cell_pointwise(const float* z, const float* bias, const float* w_i_diag,
                                      const float* w_f_diag, const float* w_o_diag,
                                      const float* c_prev, int B, int H, float forget_bias,
                                      float cell_clip, float* c_out, float* h_out,
                                      b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(z && c_prev && c_out && h_out && B > 0 && H > 0, "b2_lstm_cell_pointwise: bad argument");
  B2_CHECK_ARG((!w_i_diag && !w_f_diag && !w_o_diag) || (w_i_diag && w_f_diag && w_o_diag),
               "b2_lstm_cell_pointwise: give all three peephole vectors or none");
  lstm_cell_pointwise_kernel<<<cdiv((int64_t)B * H, 256), 256, 0, stream>>>(
      z, bias, w_i_diag, w_f_diag, w_o_diag, c_prev, B, H, forget_bias, cell_clip, c_out, h_out);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_tanh_inplace(float* x, int64_t n, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(x && n > 0, "b2_tanh_inplace: bad argument");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  tanh_kernel<<<blocks, 256, 0, stream>>>(x, n);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_argmax_rows(const float* x, int64_t rows, int C, int32_t* out, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(x && out && rows > 0 && C > 0, "b2_argmax_rows: bad argument");
  argmax_rows_kernel<<<cdiv(rows, 8), 256, 0, stream>>>(x, rows, C, out);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_decoder_step_emit(int B, int C, int Hd, int E, int T, int emb_dim, int t, int L,
                                    const float* logits, const int32_t* ids, const float* av,
                                    const float* alpha, const float* ctx, const float* c_new,
                                    const float* h_new, float* c_state, float* h_state,
                                    int32_t* finished, const float* embedding,
                                    const int32_t* labels, int labels_ld, const int32_t* dec_len,
                                    int eos, int max_iter, float* xh, float* out_logits,
                                    int32_t* out_ids, float* out_av, float* out_alpha,
                                    float* out_ctx, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(logits && ids && av && alpha && ctx && c_new && h_new && c_state && h_state &&
               finished && embedding && xh && out_logits && out_ids && out_av && out_alpha && out_ctx,
               "b2_decoder_step_emit: null pointer");
  B2_CHECK_ARG(B > 0 && C > 0 && Hd > 0 && E > 0 && T > 0 && emb_dim > 0 && t >= 0 && t < L,
               "b2_decoder_step_emit: bad shape (t=%d, L=%d)", t, L);
  B2_CHECK_ARG(!labels || (dec_len && labels_ld > 1), "b2_decoder_step_emit: teacher forcing needs dec_len");
  EmitArgs a;
  a.B = B; a.C = C; a.Hd = Hd; a.E = E; a.T = T; a.emb = emb_dim; a.t = t; a.L = L;
  a.logits = logits; a.ids = ids; a.av = av; a.alpha = alpha; a.ctx = ctx; a.c_new = c_new; a.h_new = h_new;
  a.c_state = c_state; a.h_state = h_state; a.finished = finished; a.embedding = embedding;
  a.labels = labels; a.labels_ld = labels_ld; a.dec_len = dec_len; a.eos = eos; a.max_iter = max_iter;
  a.xh = xh; a.out_logits = out_logits; a.out_ids = out_ids; a.out_av = out_av; a.out_alpha = out_alpha;
  a.out_ctx = out_ctx;
  decoder_step_emit_kernel<<<B, 256, 0, stream>>>(a);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
