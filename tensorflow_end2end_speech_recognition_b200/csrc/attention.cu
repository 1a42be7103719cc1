// Attention step (score + masked softmax + context) for sm_100a.
//
// Replaces AttentionLayer.__call__ / _compute_attention_score of the reference
// (models/attention/decoders/attention_layer.py:45-113, :115-347) for one decoder step over all
// T encoder states.  The reference recomputes W_keys.h_enc inside every step (:151-159); here
// the key projection is HOISTED (one GEMM per batch, b2_gemm) and the step kernel reads
//   keys [B,T,A]  (W_keys.h + b   | h itself for luong_dot | W_keys.h for luong_general |
//                  W_concat[:E].h for luong_concat | absent for `location`)
//   q    [B,A]    (W_query.s | s | W_concat[E:].s)
// Two energy forms cover the seven implemented types:
//   additive        e_t = sum_a v_a * tanh(keys[t,a] + q[a] + loc[t,a])      (bahdanau_content,
//                         hybrid, location, luong_concat)
//   multiplicative  e_t = sum_a keys[t,a] * q[a]                             (dot_product,
//                         luong_dot, luong_general)
// loc[t,:] = conv1d_SAME(alpha_prev, F[k,10])[t,:] . W_filter + b_filter  (location / hybrid).
// HBM-bound: per step it must read keys (4*B*T*A) and the encoder states (4*B*T*E) once.
// Two launches: attention_step_kernel (one CTA per utterance: energies, mask, normalise) and
// attention_context_kernel (B x E/256 CTAs streaming the encoder states once).
#include "common.cuh"
#include <float.h>

namespace b2 {

struct AttnArgs {
  int mode;                      // 0 additive, 1 multiplicative
  const float* enc; const float* keys; const float* q; const float* prev_alpha;
  const int* enc_len;
  const float* filt; const float* w_f; const float* b_f; const float* v_a;
  int B, T, E, A, Kw;
  float sharpening; int sigmoid_smoothing;
  float* alpha; float* context;
};

constexpr int kAttnThreads = 512;

__global__ void __launch_bounds__(kAttnThreads)
attention_step_kernel(const AttnArgs a) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kAttnThreads / 32;
  const int T = a.T, A = a.A;
  const int len = min(a.enc_len[b], T);
  float* s_e = sm;                       // [T] energies -> weights
  float* s_q = s_e + T;                  // [A]
  float* s_v = s_q + A;                  // [A]
  float* s_bf = s_v + A;                 // [A]
  float* s_wf = s_bf + A;                // [10*A]
  float* s_pa = s_wf + 10 * A;           // [T + Kw] zero-padded previous weights
  float* s_f = s_pa + (a.filt ? T + a.Kw : 0);    // [T*10] conv features
  float* s_filt = s_f + (a.filt ? T * 10 : 0);    // [Kw*10]
  __shared__ float s_red[32];

  for (int i = tid; i < A; i += kAttnThreads) {
    s_q[i] = a.q ? a.q[(size_t)b * A + i] : 0.f;
    s_v[i] = a.v_a ? a.v_a[i] : 1.f;
    s_bf[i] = a.b_f ? a.b_f[i] : 0.f;
  }
  const bool loc = a.filt != nullptr;
  if (loc) {
    const int pl = (a.Kw - 1) / 2;
    for (int i = tid; i < 10 * A; i += kAttnThreads) s_wf[i] = a.w_f[i];
    for (int i = tid; i < a.Kw * 10; i += kAttnThreads) s_filt[i] = a.filt[i];
    for (int i = tid; i < T + a.Kw; i += kAttnThreads) {
      const int t = i - pl;
      s_pa[i] = (t >= 0 && t < T) ? a.prev_alpha[(size_t)b * T + t] : 0.f;
    }
    __syncthreads();
    // conv features f[t][k] = sum_j pa[t + j - pl] * F[j][k]
    for (int i = tid; i < len * 10; i += kAttnThreads) {
      const int t = i / 10, k = i % 10;
      float acc = 0.f;
      for (int j = 0; j < a.Kw; ++j) acc = fmaf(s_pa[t + j], s_filt[j * 10 + k], acc);
      s_f[i] = acc;
    }
  }
  __syncthreads();
  // energies: one warp per t, lanes over a
  for (int t = warp; t < T; t += nwarp) {
    float e;
    if (t < len) {
      const float* kr = a.keys ? a.keys + ((size_t)b * T + t) * A : nullptr;
      float acc = 0.f;
      for (int i = lane; i < A; i += 32) {
        const float kv = kr ? kr[i] : 0.f;
        if (a.mode == 1) acc = fmaf(kv, s_q[i], acc);
        else {
          float x = kv + s_q[i];
          if (loc) {
            float l = s_bf[i];
#pragma unroll
            for (int k = 0; k < 10; ++k) l = fmaf(s_f[t * 10 + k], s_wf[k * A + i], l);
            x += l;
          }
          acc = fmaf(s_v[i], tanhf_(x), acc);
        }
      }
      e = warp_sum(acc);
    } else {
      e = -FLT_MAX;                                   // tf.float32.min (attention_layer.py:84-85)
    }
    if (lane == 0) s_e[t] = e * a.sharpening;
  }
  __syncthreads();
  // normalise over T
  float m = -INFINITY;
  if (!a.sigmoid_smoothing) {
    for (int t = tid; t < T; t += kAttnThreads) m = fmaxf(m, s_e[t]);
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = (lane < nwarp) ? s_red[lane] : -INFINITY;
    m = warp_max(m);
    __syncthreads();
  }
  float sum = 0.f;
  for (int t = tid; t < T; t += kAttnThreads) {
    const float e = s_e[t];
    float w;
    if (a.sigmoid_smoothing) w = (t < len) ? 1.f / (1.f + __expf(-e)) : 0.f;
    else w = __expf(e - m);
    s_e[t] = w;
    sum += w;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = (lane < nwarp) ? s_red[lane] : 0.f;
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int t = tid; t < T; t += kAttnThreads) {
    const float w = s_e[t] * inv;
    a.alpha[(size_t)b * T + t] = w;
  }
}

// context[b, e] = sum_{t < len} alpha[b,t] * enc[b,t,e].  CTA = 64 float4 columns (256
// features) x 8 time groups; weights past `len` are exactly 0 so padded frames are never read.
__global__ void __launch_bounds__(512)
attention_context_kernel(const float* __restrict__ enc, const float* __restrict__ alpha,
                         const int* __restrict__ enc_len, int T, int E, float* __restrict__ context) {
  __shared__ float4 red[8][64];
  const int b = blockIdx.x;
  const int col = blockIdx.y * 64 + (threadIdx.x & 63);      // float4 column
  const int tg = threadIdx.x >> 6;
  const int len = min(enc_len[b], T);
  const int E4 = E / 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < E4) {
    const float4* encb = (const float4*)(enc + (size_t)b * T * E) + col;
    const float* al = alpha + (size_t)b * T;
    int t = tg;
    for (; t + 24 < len; t += 32) {
      const float4 h0 = __ldg(encb + (size_t)t * E4), h1 = __ldg(encb + (size_t)(t + 8) * E4);
      const float4 h2 = __ldg(encb + (size_t)(t + 16) * E4), h3 = __ldg(encb + (size_t)(t + 24) * E4);
      const float w0 = al[t], w1 = al[t + 8], w2 = al[t + 16], w3 = al[t + 24];
      acc.x += w0 * h0.x + w1 * h1.x + w2 * h2.x + w3 * h3.x;
      acc.y += w0 * h0.y + w1 * h1.y + w2 * h2.y + w3 * h3.y;
      acc.z += w0 * h0.z + w1 * h1.z + w2 * h2.z + w3 * h3.z;
      acc.w += w0 * h0.w + w1 * h1.w + w2 * h2.w + w3 * h3.w;
    }
    for (; t < len; t += 8) {
      const float4 h = __ldg(encb + (size_t)t * E4);
      const float w = al[t];
      acc.x += w * h.x; acc.y += w * h.y; acc.z += w * h.z; acc.w += w * h.w;
    }
  }
  red[tg][threadIdx.x & 63] = acc;
  __syncthreads();
  if (tg == 0 && col < E4) {
#pragma unroll
    for (int g = 1; g < 8; ++g) {
      const float4 o = red[g][threadIdx.x & 63];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    ((float4*)(context + (size_t)b * E))[col] = acc;
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_attention_step_forward(int mode, const float* enc, const float* keys, const float* q,
                                         const float* prev_alpha, const int32_t* enc_len,
                                         const float* conv_filter, int filter_width,
                                         const float* w_filter, const float* b_filter,
                                         const float* v_a, int B, int T, int E, int A,
                                         float sharpening_factor, int sigmoid_smoothing,
                                         float* alpha, float* context, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(enc && enc_len && alpha && context, "b2_attention_step_forward: null pointer");
  B2_CHECK_ARG(mode == 0 || mode == 1, "b2_attention_step_forward: mode %d", mode);
  B2_CHECK_ARG(B > 0 && T > 0 && E > 0 && A > 0 && E % 4 == 0, "b2_attention_step_forward: bad shape");
  B2_CHECK_ARG(keys || (mode == 0 && conv_filter), "b2_attention_step_forward: keys missing");
  B2_CHECK_ARG(!conv_filter || (prev_alpha && w_filter && filter_width > 0),
               "b2_attention_step_forward: location term needs prev_alpha / W_filter");
  AttnArgs a;
  a.mode = mode; a.enc = enc; a.keys = keys; a.q = q; a.prev_alpha = prev_alpha; a.enc_len = enc_len;
  a.filt = conv_filter; a.w_f = w_filter; a.b_f = b_filter; a.v_a = v_a;
  a.B = B; a.T = T; a.E = E; a.A = A; a.Kw = conv_filter ? filter_width : 0;
  a.sharpening = sharpening_factor; a.sigmoid_smoothing = sigmoid_smoothing;
  a.alpha = alpha; a.context = context;
  size_t smem = ((size_t)T + 3 * A + 10 * A) * 4;
  if (conv_filter) smem += ((size_t)T + filter_width + (size_t)T * 10 + (size_t)filter_width * 10) * 4;
  B2_CHECK_ARG(smem <= 200 * 1024, "b2_attention_step_forward: T=%d too long for shared memory", T);
  B2_CUDA(cudaFuncSetAttribute(attention_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_step_kernel<<<B, kAttnThreads, smem, stream>>>(a);
  B2_LAUNCH_CHECK();
  dim3 cgrid(B, cdiv(E / 4, 64));
  attention_context_kernel<<<cgrid, 512, 0, stream>>>(enc, alpha, enc_len, T, E, context);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
