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
// Three launches: attention_energy_kernel (B x T/64 CTAs: energies + mask), attention_normalise_kernel
// (one CTA per row, in place) and attention_context_kernel (B x E/256 CTAs streaming the encoder
// states once; one skinny GEMM per utterance when beam rows share them).
#include "common.cuh"
#include <float.h>

namespace b2 {
int num_sms();
int gemm_skinny_batched(int M, int N, int K, const float* A, int lda, int64_t strideA, const float* B, int ldb,
                        int64_t strideB, float* C, int ldc, int64_t strideC, int batch, cudaStream_t stream);
}

namespace b2 {

struct AttnArgs {
  int mode;                      // 0 additive, 1 multiplicative
  const float* enc; const float* keys; const float* q; const float* prev_alpha;
  const int* enc_len;
  const float* filt; const float* w_f; const float* b_f; const float* v_a;
  int B, T, E, A, Kw;
  float sharpening; int sigmoid_smoothing;
  float* alpha; float* context; float* energy;
  int rpu;                       // batch rows per utterance (beam search: beam_width rows share one enc/keys)
};

constexpr int kAttnThreads = 256;
constexpr int kEnergyChunk = 64;        // frames per CTA of the energy kernel

// energies of kEnergyChunk frames of one batch row: grid (rows, ceil(T / chunk)).  Written (sharpened,
// padded frames = float32.min * sharpening) into the `alpha` buffer, normalised in place afterwards.
__global__ void __launch_bounds__(kAttnThreads)
attention_energy_kernel(const AttnArgs a) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int ub = b / a.rpu;             // utterance whose encoder states this row attends to
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kAttnThreads / 32;
  const int T = a.T, A = a.A;
  const int len = min(a.enc_len[ub], T);
  const int t0 = blockIdx.y * kEnergyChunk;
  const int tn = min(kEnergyChunk, T - t0);
  float* s_q = sm;                       // [A]
  float* s_v = s_q + A;                  // [A]
  float* s_bf = s_v + A;                 // [A]
  float* s_wf = s_bf + A;                // [10*A]            (location term with real previous weights)
  float* s_pa = s_wf + 10 * A;           // [chunk + Kw]      zero-padded previous weights
  float* s_f = s_pa + kEnergyChunk + a.Kw;        // [chunk*10] conv features
  float* s_filt = s_f + kEnergyChunk * 10;        // [Kw*10]
  for (int i = tid; i < A; i += kAttnThreads) {
    s_q[i] = a.q ? a.q[(size_t)b * A + i] : 0.f;
    s_v[i] = a.v_a ? a.v_a[i] : 1.f;
    s_bf[i] = a.b_f ? a.b_f[i] : 0.f;
  }
  const bool loc = a.filt != nullptr;
  // prev_alpha == NULL with a location term: the previous weights are known to be all zero (what the
  // reference's decoder always feeds) -> conv features are 0, the term is b_filter
  const bool loc_conv = loc && a.prev_alpha != nullptr;
  if (loc_conv && t0 < len) {
    const int pl = (a.Kw - 1) / 2;
    for (int i = tid; i < 10 * A; i += kAttnThreads) s_wf[i] = a.w_f[i];
    for (int i = tid; i < a.Kw * 10; i += kAttnThreads) s_filt[i] = a.filt[i];
    for (int i = tid; i < kEnergyChunk + a.Kw; i += kAttnThreads) {
      const int t = t0 + i - pl;
      s_pa[i] = (t >= 0 && t < T) ? a.prev_alpha[(size_t)b * T + t] : 0.f;
    }
    __syncthreads();
    // conv features f[t][k] = sum_j pa[t + j - pl] * F[j][k]; thread = one frame, 10 accumulators
    for (int tl = tid; tl < tn; tl += kAttnThreads) {
      float acc[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) acc[k] = 0.f;
      for (int j = 0; j < a.Kw; ++j) {
        const float pv = s_pa[tl + j];
        const float* fr = s_filt + j * 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[k] = fmaf(pv, fr[k], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < 10; ++k) s_f[tl * 10 + k] = acc[k];
    }
  }
  __syncthreads();
  for (int tl = warp; tl < tn; tl += nwarp) {
    const int t = t0 + tl;
    float e;
    if (t < len) {
      const float* kr = a.keys ? a.keys + ((size_t)ub * T + t) * A : nullptr;
      float acc = 0.f;
      for (int i = lane; i < A; i += 32) {
        const float kv = kr ? kr[i] : 0.f;
        if (a.mode == 1) acc = fmaf(kv, s_q[i], acc);
        else {
          float x = kv + s_q[i];
          if (loc_conv) {
            float l = s_bf[i];
#pragma unroll
            for (int k = 0; k < 10; ++k) l = fmaf(s_f[tl * 10 + k], s_wf[k * A + i], l);
            x += l;
          } else if (loc) {
            x += s_bf[i];
          }
          acc = fmaf(s_v[i], tanhf_(x), acc);
        }
      }
      e = warp_sum(acc);
    } else {
      e = -FLT_MAX;                                   // tf.float32.min (attention_layer.py:84-85)
    }
    if (lane == 0) a.alpha[(size_t)b * T + t] = e * a.sharpening;
  }
}

// one CTA per batch row: energies (in `alpha`) -> softmax or sigmoid / sum, in place
__global__ void __launch_bounds__(512)
attention_normalise_kernel(const AttnArgs a) {
  extern __shared__ float s_e[];         // [T]
  __shared__ float s_red[32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = 16;
  const int T = a.T;
  const int len = min(a.enc_len[b / a.rpu], T);
  float* al = a.alpha + (size_t)b * T;
  for (int t = tid; t < T; t += 512) {
    const float e = al[t];
    s_e[t] = e;
    if (a.energy && t < len) a.energy[(size_t)b * T + t] = e;
  }
  __syncthreads();
  float m = -INFINITY;
  if (!a.sigmoid_smoothing) {
    for (int t = tid; t < T; t += 512) m = fmaxf(m, s_e[t]);
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = (lane < nwarp) ? s_red[lane] : -INFINITY;
    m = warp_max(m);
    __syncthreads();
  }
  float sum = 0.f;
  for (int t = tid; t < T; t += 512) {
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
  for (int t = tid; t < T; t += 512) al[t] = s_e[t] * inv;
}

// context[b, e] = sum_{t < len} alpha[b,t] * enc[b,t,e].  CTA = 64 float4 columns (256
// features) x 8 time groups; weights past `len` are exactly 0 so padded frames are never read.
__global__ void __launch_bounds__(512)
attention_context_kernel(const float* __restrict__ enc, const float* __restrict__ alpha,
                         const int* __restrict__ enc_len, int T, int E, int rpu, float* __restrict__ context) {
  __shared__ float4 red[8][64];
  const int b = blockIdx.x;
  const int ub = b / rpu;
  const int col = blockIdx.y * 64 + (threadIdx.x & 63);      // float4 column
  const int tg = threadIdx.x >> 6;
  // blockIdx.z: slice of the time axis (few rows -> more CTAs; partial sums are added atomically into
  // a zeroed context)
  const int nz = gridDim.z;
  const int tchunk = ((T + nz - 1) / nz + 7) / 8 * 8;
  const int tbeg = blockIdx.z * tchunk;
  const int len = min(min(enc_len[ub], T), tbeg + tchunk);
  const int E4 = E / 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < E4) {
    const float4* encb = (const float4*)(enc + (size_t)ub * T * E) + col;
    const float* al = alpha + (size_t)b * T;
    int t = tbeg + tg;
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
    float* cp = context + (size_t)b * E + col * 4;
    if (nz == 1) *(float4*)cp = acc;
    else if (tbeg < len) { atomicAdd(cp, acc.x); atomicAdd(cp + 1, acc.y); atomicAdd(cp + 2, acc.z); atomicAdd(cp + 3, acc.w); }
  }
}

// ---------------------------------------------------------------------------------------
// Backward of one attention step (training).  Given d(context) of this decoder step:
//   K1  dalpha[b,t] = enc[b,t,:] . dctx[b,:]   and the row reductions sum_t alpha*dalpha
//       (softmax / sigmoid-normalisation backward) and, for sigmoid smoothing, S = sum_t sig(e)
//   K2  de[t] -> energy backward: d_keys += ..., dq, dv, db_filter
// d(enc) through the context (alpha_t (x) dctx_t) is a rank-L update per utterance and is left
// to one GEMM per utterance after the decoder loop; here only the per-step sequential part.
// The location term is the constant b_filter (the reference feeds zero previous weights,
// SURVEY A.7.1), so the filter and W_filter get zero gradient.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attention_bwd_dalpha_kernel(const float* __restrict__ enc, const float* __restrict__ dctx,
                            const float* __restrict__ alpha, const float* __restrict__ energy,
                            const int* __restrict__ enc_len, int T, int E, int sigmoid_smoothing,
                            const float* __restrict__ dalpha_ext, float* __restrict__ dalpha,
                            float* __restrict__ red) {
  extern __shared__ float s_d[];                    // [E]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = min(enc_len[b], T);
  for (int i = threadIdx.x; i < E; i += 256) s_d[i] = dctx[(size_t)b * E + i];
  __syncthreads();
  float dot = 0.f, ssum = 0.f;
  const int t0 = blockIdx.y * 64;
  for (int t = t0 + warp; t < min(t0 + 64, T); t += 8) {
    float acc = 0.f;
    if (t < len) {
      const float4* er = (const float4*)(enc + ((size_t)b * T + t) * E);
      for (int i = lane; i < E / 4; i += 32) {
        const float4 h = __ldg(er + i);
        const float4 d = ((const float4*)s_d)[i];
        acc += h.x * d.x + h.y * d.y + h.z * d.z + h.w * d.w;
      }
      acc = warp_sum(acc);
      if (dalpha_ext) acc += dalpha_ext[(size_t)b * T + t];   // gradient through the next step's location term
      if (lane == 0) {
        dot += alpha[(size_t)b * T + t] * acc;
        if (sigmoid_smoothing) ssum += 1.f / (1.f + __expf(-energy[(size_t)b * T + t]));
      }
    }
    if (lane == 0) dalpha[(size_t)b * T + t] = acc;
  }
  if (lane == 0) {
    if (dot != 0.f) atomicAdd(&red[b * 2], dot);
    if (ssum != 0.f) atomicAdd(&red[b * 2 + 1], ssum);
  }
}

struct AttnBwdArgs {
  int mode;
  const float* keys; const float* q; const float* alpha; const float* energy; const float* dalpha;
  const float* red; const int* enc_len; const float* b_f; const float* v_a;
  int T, A; float sharpening; int sigmoid_smoothing;
  float* d_keys; float* dq; float* dv; float* db_f;
  // location term with REAL previous weights (feed_previous_attention): recompute f = conv(prev_alpha, F),
  // accumulate d(W_filter), emit df [B,T,10] for attention_bwd_location_kernel
  const float* prev_alpha; const float* filt; const float* w_f; int Kw;
  float* d_w_f; float* df;
};

// grid (B, ceil(T/32)); 8 warps, warp w takes rows t0+w, t0+w+8, ...
__global__ void __launch_bounds__(256) attention_bwd_energy_kernel(const AttnBwdArgs a) {
  extern __shared__ float sm[];
  const int A = a.A, T = a.T;
  float* s_q = sm;                 // [A]
  float* s_v = s_q + A;            // [A]
  float* s_b = s_v + A;            // [A]
  float* s_dq = s_b + A;           // [8][A]
  float* s_dv = s_dq + 8 * A;      // [8][A]
  // the next five only with prev_alpha:
  float* s_wf = s_dv + 8 * A;      // [10][A]
  float* s_dwf = s_wf + 10 * A;    // [10][A]
  float* s_pa = s_dwf + 10 * A;    // [32 + Kw]
  float* s_f = s_pa + 32 + a.Kw;   // [32][10]
  float* s_filt = s_f + 320;       // [Kw][10]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = min(a.enc_len[b], T);
  const bool conv = a.prev_alpha != nullptr;
  for (int i = threadIdx.x; i < A; i += 256) {
    s_q[i] = a.q ? a.q[(size_t)b * A + i] : 0.f;
    s_v[i] = a.v_a ? a.v_a[i] : 1.f;
    s_b[i] = a.b_f ? a.b_f[i] : 0.f;
  }
  for (int i = threadIdx.x; i < 16 * A; i += 256) s_dq[i] = 0.f;
  const int t0 = blockIdx.y * 32;
  if (conv) {
    const int pl = (a.Kw - 1) / 2;
    for (int i = threadIdx.x; i < 10 * A; i += 256) { s_wf[i] = a.w_f[i]; s_dwf[i] = 0.f; }
    for (int i = threadIdx.x; i < a.Kw * 10; i += 256) s_filt[i] = a.filt[i];
    for (int i = threadIdx.x; i < 32 + a.Kw; i += 256) {
      const int t = t0 + i - pl;
      s_pa[i] = (t >= 0 && t < T) ? a.prev_alpha[(size_t)b * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 320; i += 256) {
      const int tl = i / 10, k = i % 10;
      float acc = 0.f;
      for (int j = 0; j < a.Kw; ++j) acc = fmaf(s_pa[tl + j], s_filt[j * 10 + k], acc);
      s_f[i] = acc;
    }
  }
  __syncthreads();
  const float dot = a.red[b * 2];
  const float S = a.sigmoid_smoothing ? a.red[b * 2 + 1] : 1.f;
  float* my_dq = s_dq + warp * A;
  float* my_dv = s_dv + warp * A;
  for (int t = t0 + warp; t < min(t0 + 32, len); t += 8) {
    const size_t bt = (size_t)b * T + t;
    float de;
    if (a.sigmoid_smoothing) {
      const float s = 1.f / (1.f + __expf(-a.energy[bt]));
      de = (a.dalpha[bt] - dot) / S * s * (1.f - s) * a.sharpening;
    } else {
      de = a.alpha[bt] * (a.dalpha[bt] - dot) * a.sharpening;
    }
    const float* kr = a.keys ? a.keys + bt * A : nullptr;
    float* dk = a.d_keys ? a.d_keys + bt * A : nullptr;
    const float* fr = s_f + (t - t0) * 10;
    float dfk[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) dfk[k] = 0.f;
    for (int i = lane; i < A; i += 32) {
      const float kv = kr ? kr[i] : 0.f;
      if (a.mode == 1) {
        if (dk) dk[i] += de * s_q[i];
        my_dq[i] += de * kv;
      } else {
        float l = s_b[i];
        if (conv) {
#pragma unroll
          for (int k = 0; k < 10; ++k) l = fmaf(fr[k], s_wf[k * A + i], l);
        }
        const float u = tanhf_(kv + s_q[i] + l);
        const float g = de * s_v[i] * (1.f - u * u);
        if (dk) dk[i] += g;
        my_dq[i] += g;
        my_dv[i] += de * u;
        if (conv) {
#pragma unroll
          for (int k = 0; k < 10; ++k) {
            dfk[k] = fmaf(g, s_wf[k * A + i], dfk[k]);
            atomicAdd(&s_dwf[k * A + i], fr[k] * g);       // d(W_filter)[k,a] += f[t,k] * g[t,a]
          }
        }
      }
    }
    if (conv) {
#pragma unroll
      for (int k = 0; k < 10; ++k) dfk[k] = warp_sum(dfk[k]);
      if (lane < 10) {
        float v = dfk[0];
#pragma unroll
        for (int k = 1; k < 10; ++k) v = (lane == k) ? dfk[k] : v;
        a.df[bt * 10 + lane] = v;
      }
    }
  }
  if (conv) {      // frames of this chunk past the utterance's end (or never visited) carry no gradient
    for (int i = threadIdx.x; i < 320; i += 256) {
      const int t = t0 + i / 10;
      if (t < T && t >= len) a.df[((size_t)b * T + t) * 10 + i % 10] = 0.f;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A; i += 256) {
    float sq = 0.f, sv = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { sq += s_dq[w * A + i]; sv += s_dv[w * A + i]; }
    if (sq != 0.f) {
      atomicAdd(&a.dq[(size_t)b * A + i], sq);
      if (a.db_f) atomicAdd(&a.db_f[i], sq);
    }
    if (a.dv && sv != 0.f) atomicAdd(&a.dv[i], sv);
  }
  if (conv)
    for (int i = threadIdx.x; i < 10 * A; i += 256)
      if (s_dwf[i] != 0.f) atomicAdd(&a.d_w_f[i], s_dwf[i]);
}

// location term, second half: from df [B,T,10] (gradient wrt the conv features)
//   d_prev_alpha[b,s] = sum_{j,k} F[j,k] * df[b, s - j + pl, k]
//   dF[j,k]          += sum_{b,t} prev_alpha[b, t + j - pl] * df[b,t,k]
// grid (B, ceil(T/64)), 256 threads
__global__ void __launch_bounds__(256)
attention_bwd_location_kernel(const float* __restrict__ df, const float* __restrict__ prev_alpha,
                              const float* __restrict__ filt, int T, int Kw, float* __restrict__ d_prev_alpha,
                              float* __restrict__ d_filt) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, t0 = blockIdx.y * 64, pl = (Kw - 1) / 2;
  float* s_filt = sm;                       // [Kw][10]
  float* s_df = s_filt + Kw * 10;           // [64 + Kw][10]: frames t0 - (Kw-1-pl) .. t0 + 63 + pl
  float* s_pa = s_df + (64 + Kw) * 10;      // [64 + Kw]:     frames t0 - pl .. t0 + 63 + (Kw-1-pl)
  const int lo = t0 - (Kw - 1 - pl);
  for (int i = threadIdx.x; i < Kw * 10; i += 256) s_filt[i] = filt[i];
  for (int i = threadIdx.x; i < (64 + Kw) * 10; i += 256) {
    const int t = lo + i / 10;
    s_df[i] = (t >= 0 && t < T) ? df[((size_t)b * T + t) * 10 + i % 10] : 0.f;
  }
  for (int i = threadIdx.x; i < 64 + Kw; i += 256) {
    const int t = t0 - pl + i;
    s_pa[i] = (t >= 0 && t < T) ? prev_alpha[(size_t)b * T + t] : 0.f;
  }
  __syncthreads();
  // d_prev_alpha for s = t0 .. t0+63: frame index s - j + pl -> local (s - j + pl) - lo
  for (int sl = threadIdx.x; sl < 64; sl += 256) {
    const int s = t0 + sl;
    if (s >= T) continue;
    float acc = 0.f;
    for (int j = 0; j < Kw; ++j) {
      const float* dr = s_df + (s - j + pl - lo) * 10;
      const float* fr = s_filt + j * 10;
#pragma unroll
      for (int k = 0; k < 10; ++k) acc = fmaf(fr[k], dr[k], acc);
    }
    d_prev_alpha[(size_t)b * T + s] = acc;
  }
  // dF[j,k] += sum_{t in chunk} pa[t + j - pl] * df[t,k]
  for (int i = threadIdx.x; i < Kw * 10; i += 256) {
    const int j = i / 10, k = i % 10;
    float acc = 0.f;
    for (int tl = 0; tl < 64; ++tl) {
      if (t0 + tl >= T) break;
      acc = fmaf(s_pa[tl + j], s_df[(t0 + tl - lo) * 10 + k], acc);
    }
    if (acc != 0.f) atomicAdd(&d_filt[i], acc);
  }
}

}  // namespace b2

using namespace b2;

namespace b2 {
int attention_step_forward_rows(int mode, const float* enc, const float* keys, const float* q,
                                const float* prev_alpha, const int32_t* enc_len,
                                const float* conv_filter, int filter_width,
                                const float* w_filter, const float* b_filter,
                                const float* v_a, int B, int T, int E, int A,
                                float sharpening_factor, int sigmoid_smoothing,
                                float* alpha, float* context, float* energy_out, int rows_per_utt,
                                b2_stream_t stream_);
}

extern "C" int b2_attention_step_forward(int mode, const float* enc, const float* keys, const float* q,
                                         const float* prev_alpha, const int32_t* enc_len,
                                         const float* conv_filter, int filter_width,
                                         const float* w_filter, const float* b_filter,
                                         const float* v_a, int B, int T, int E, int A,
                                         float sharpening_factor, int sigmoid_smoothing,
                                         float* alpha, float* context, float* energy_out,
                                         b2_stream_t stream_) {
  return attention_step_forward_rows(mode, enc, keys, q, prev_alpha, enc_len, conv_filter, filter_width, w_filter,
                                     b_filter, v_a, B, T, E, A, sharpening_factor, sigmoid_smoothing, alpha,
                                     context, energy_out, 1, stream_);
}

// B batch rows; rows_per_utt consecutive rows share one utterance's enc / keys / enc_len
int b2::attention_step_forward_rows(int mode, const float* enc, const float* keys, const float* q,
                                    const float* prev_alpha, const int32_t* enc_len,
                                    const float* conv_filter, int filter_width,
                                    const float* w_filter, const float* b_filter,
                                    const float* v_a, int B, int T, int E, int A,
                                    float sharpening_factor, int sigmoid_smoothing,
                                    float* alpha, float* context, float* energy_out, int rows_per_utt,
                                    b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(enc && enc_len && alpha && context, "b2_attention_step_forward: null pointer");
  B2_CHECK_ARG(mode == 0 || mode == 1, "b2_attention_step_forward: mode %d", mode);
  B2_CHECK_ARG(B > 0 && T > 0 && E > 0 && A > 0 && E % 4 == 0, "b2_attention_step_forward: bad shape");
  B2_CHECK_ARG(keys || (mode == 0 && conv_filter), "b2_attention_step_forward: keys missing");
  B2_CHECK_ARG(!conv_filter || (w_filter && filter_width > 0),
               "b2_attention_step_forward: location term needs W_filter");
  AttnArgs a;
  a.mode = mode; a.enc = enc; a.keys = keys; a.q = q; a.prev_alpha = prev_alpha; a.enc_len = enc_len;
  a.filt = conv_filter; a.w_f = w_filter; a.b_f = b_filter; a.v_a = v_a;
  a.B = B; a.T = T; a.E = E; a.A = A; a.Kw = conv_filter ? filter_width : 0;
  a.sharpening = sharpening_factor; a.sigmoid_smoothing = sigmoid_smoothing;
  a.alpha = alpha; a.context = context; a.energy = energy_out; a.rpu = rows_per_utt > 0 ? rows_per_utt : 1;
  size_t smem = ((size_t)13 * A + kEnergyChunk + a.Kw + kEnergyChunk * 10 + (size_t)a.Kw * 10) * 4;
  B2_CHECK_ARG(smem <= 200 * 1024 && (size_t)T * 4 <= 200 * 1024,
               "b2_attention_step_forward: A=%d / T=%d too large for shared memory", A, T);
  B2_CUDA(cudaFuncSetAttribute(attention_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B2_CUDA(cudaFuncSetAttribute(attention_normalise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 4));
  dim3 egrid(B, cdiv(T, kEnergyChunk));
  attention_energy_kernel<<<egrid, kAttnThreads, smem, stream>>>(a);
  B2_LAUNCH_CHECK();
  attention_normalise_kernel<<<B, 512, (size_t)T * 4, stream>>>(a);
  B2_LAUNCH_CHECK();
  dim3 cgrid(B, cdiv(E / 4, 64));
  B2_CHECK_ARG(a.rpu <= 64, "attention_step_forward_rows: at most 64 rows per utterance");
  if (a.rpu > 1) {
    // beam rows of one utterance share its encoder states: context[u] = Alpha[u] [W,T] . enc[u] [T,E], one
    // skinny product per utterance, enc streamed once (weights past enc_len are exactly 0)
    int rc = gemm_skinny_batched(a.rpu, E, T, alpha, T, (int64_t)a.rpu * T, enc, E, (int64_t)T * E, context, E,
                                 (int64_t)a.rpu * E, B / a.rpu, stream);
    if (rc) return rc;
  } else {
    // enough CTAs to fill the machine: split the time axis when the batch is small
    int nz = 1;
    while ((int64_t)cgrid.x * cgrid.y * nz < 2 * num_sms() && nz < 16 && T / (nz * 2) >= 64) nz *= 2;
    cgrid.z = nz;
    if (nz > 1) B2_CUDA(cudaMemsetAsync(context, 0, (size_t)B * E * sizeof(float), stream));
    attention_context_kernel<<<cgrid, 512, 0, stream>>>(enc, alpha, enc_len, T, E, a.rpu, context);
  }
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" size_t b2_attention_step_backward_workspace_bytes(int B, int T) {
  return align_up((size_t)B * T * 4, 256) + align_up((size_t)B * 2 * 4, 256) + align_up((size_t)B * T * 10 * 4, 256);
}

extern "C" int b2_attention_step_backward(int mode, const float* enc, const float* keys, const float* q,
                                          const float* alpha, const float* energy,
                                          const int32_t* enc_len, const float* b_filter,
                                          const float* v_a, int B, int T, int E, int A,
                                          float sharpening_factor, int sigmoid_smoothing,
                                          const float* dctx, float* d_keys, float* dq,
                                          int dq_accumulate, float* dv,
                                          float* db_filter, void* workspace, size_t workspace_bytes,
                                          b2_stream_t stream_) {
  return b2_attention_step_backward_loc(mode, enc, keys, q, alpha, energy, enc_len, b_filter, v_a, B, T, E, A,
                                        sharpening_factor, sigmoid_smoothing, dctx, d_keys, dq, dq_accumulate, dv,
                                        db_filter, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                        workspace, workspace_bytes, stream_);
}

extern "C" int b2_attention_step_backward_loc(int mode, const float* enc, const float* keys, const float* q,
                                              const float* alpha, const float* energy,
                                              const int32_t* enc_len, const float* b_filter,
                                              const float* v_a, int B, int T, int E, int A,
                                              float sharpening_factor, int sigmoid_smoothing,
                                              const float* dctx, float* d_keys, float* dq,
                                              int dq_accumulate, float* dv, float* db_filter,
                                              const float* prev_alpha, const float* conv_filter, int filter_width,
                                              const float* w_filter, const float* dalpha_ext,
                                              float* d_prev_alpha, float* d_conv_filter, float* d_w_filter,
                                              void* workspace, size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(enc && alpha && enc_len && dctx && dq && workspace, "b2_attention_step_backward: null pointer");
  B2_CHECK_ARG(mode == 0 || mode == 1, "b2_attention_step_backward: mode %d", mode);
  B2_CHECK_ARG(B > 0 && T > 0 && E > 0 && A > 0 && E % 4 == 0, "b2_attention_step_backward: bad shape");
  B2_CHECK_ARG(!sigmoid_smoothing || energy, "b2_attention_step_backward: sigmoid smoothing needs the saved energies");
  B2_CHECK_ARG(mode == 0 || keys, "b2_attention_step_backward: multiplicative mode needs keys");
  const bool conv = prev_alpha != nullptr;
  B2_CHECK_ARG(!conv || (mode == 0 && conv_filter && w_filter && filter_width > 0 && d_prev_alpha && d_conv_filter &&
                         d_w_filter && b_filter),
               "b2_attention_step_backward: location term with previous weights needs filter / W_filter and their gradients");
  const size_t need = b2_attention_step_backward_workspace_bytes(B, T);
  if (workspace_bytes < need) { set_error("b2_attention_step_backward: workspace %zu < %zu", workspace_bytes, need); return B2_ERR_WORKSPACE; }
  float* dalpha = (float*)workspace;
  float* red = (float*)((char*)workspace + align_up((size_t)B * T * 4, 256));
  float* df = (float*)((char*)red + align_up((size_t)B * 2 * 4, 256));
  B2_CUDA(cudaMemsetAsync(red, 0, (size_t)B * 2 * 4, stream));
  if (!dq_accumulate) B2_CUDA(cudaMemsetAsync(dq, 0, (size_t)B * A * 4, stream));
  dim3 g1(B, cdiv(T, 64));
  attention_bwd_dalpha_kernel<<<g1, 256, (size_t)E * 4, stream>>>(enc, dctx, alpha, energy, enc_len, T, E,
                                                                 sigmoid_smoothing, dalpha_ext, dalpha, red);
  B2_LAUNCH_CHECK();
  AttnBwdArgs a;
  a.mode = mode; a.keys = keys; a.q = q; a.alpha = alpha; a.energy = energy; a.dalpha = dalpha; a.red = red;
  a.enc_len = enc_len; a.b_f = b_filter; a.v_a = v_a; a.T = T; a.A = A; a.sharpening = sharpening_factor;
  a.sigmoid_smoothing = sigmoid_smoothing; a.d_keys = d_keys; a.dq = dq; a.dv = dv; a.db_f = db_filter;
  a.prev_alpha = prev_alpha; a.filt = conv_filter; a.w_f = w_filter; a.Kw = conv ? filter_width : 0;
  a.d_w_f = d_w_filter; a.df = df;
  size_t smem = (size_t)19 * A * 4;
  if (conv) smem += ((size_t)20 * A + 32 + filter_width + 320 + (size_t)filter_width * 10) * 4;
  B2_CHECK_ARG(smem <= 200 * 1024, "b2_attention_step_backward: A=%d too wide for shared memory", A);
  B2_CUDA(cudaFuncSetAttribute(attention_bwd_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 g2(B, cdiv(T, 32));
  attention_bwd_energy_kernel<<<g2, 256, smem, stream>>>(a);
  B2_LAUNCH_CHECK();
  if (conv) {
    const size_t smem3 = ((size_t)filter_width * 10 + (size_t)(64 + filter_width) * 11) * 4;
    B2_CUDA(cudaFuncSetAttribute(attention_bwd_location_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    dim3 g3(B, cdiv(T, 64));
    attention_bwd_location_kernel<<<g3, 256, smem3, stream>>>(df, prev_alpha, conv_filter, T, filter_width,
                                                             d_prev_alpha, d_conv_filter);
    B2_LAUNCH_CHECK();
  }
  return B2_OK;
}
