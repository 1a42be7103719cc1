// CTC greedy decoder + posteriors softmax for sm_100a.
//
// b2_ctc_greedy_decode replaces tf.nn.ctc_greedy_decoder (models/ctc/ctc.py:340-342;
// numpy twin models/ctc/decoders/greedy_decoder.py:19-50): per-frame argmax
// (first index on ties) over t < seq_len, collapse repeats, drop blanks.
// b2_softmax_rows replaces tf.nn.softmax in CTC.posteriors (ctc.py:354-380).
// Both are HBM-bound streaming kernels: 4*T*B*C bytes read, O(T*B) written.
#include "common.cuh"

namespace b2 {

// One CTA per utterance.  Phase 1: warp-per-frame argmax into out_labels (used
// as scratch).  Phase 2: keep-flags + block-wide exclusive scan + compaction.
__global__ void __launch_bounds__(256)
ctc_greedy_kernel(const float* __restrict__ logits, const int* __restrict__ seq_len, int T,
                  int B, int C, int blank, int* __restrict__ out_labels,
                  int* __restrict__ out_len) {
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  __shared__ int s_last;     // argmax of the frame just before the current chunk
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tb = min(seq_len[b], T);
  // The output row doubles as argmax scratch: compaction writes out[pos] with pos <= t
  // only after every read of the chunk (and the chunk's last argmax) has been taken.
  int* am = out_labels + (int64_t)b * T;
  for (int t = warp; t < Tb; t += 8) {
    const float* x = logits + ((int64_t)t * B + b) * C;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = x[c];
      if (v > best) { best = v; bi = c; }   // strict >: first index wins inside a lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;   // all -inf/NaN row: numpy argmax returns 0
    if (lane == 0) am[t] = bi;
  }
  if (tid == 0) { s_carry = 0; s_last = -1; }
  __syncthreads();
  int* out = am;
  for (int base = 0; base < Tb; base += 256) {
    const int t = base + tid;
    int k = -1, keep = 0;
    if (t < Tb) {
      k = am[t];
      const int prev = (tid > 0) ? am[t - 1] : s_last;
      keep = (k != blank && k != prev) ? 1 : 0;
    }
    // block exclusive scan of keep
    int v = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp[w];
    const int carry = s_carry;
    const int pos = carry + woff + v - keep;
    const int last = (base + 255 < Tb) ? am[base + 255] : -1;
    __syncthreads();
    if (keep) out[pos] = k;
    if (tid == 255) { s_carry = carry + woff + v; s_last = last; }
    __syncthreads();
  }
  const int n = s_carry;
  for (int t = n + tid; t < T; t += 256) out[t] = -1;
  if (tid == 0) out_len[b] = n;
}

__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float* xr = x + row * C;
  float* yr = y + row * C;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(xr[c] - m);
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int c = lane; c < C; c += 32) yr[c] = __expf(xr[c] - m) * inv;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_ctc_greedy_decode(const float* logits, const int32_t* seq_len, int T, int B,
                                    int C, int blank, int32_t* out_labels, int32_t* out_len,
                                    b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(logits && seq_len && out_labels && out_len, "b2_ctc_greedy_decode: null pointer");
  B2_CHECK_ARG(T > 0 && B > 0 && C > 0, "b2_ctc_greedy_decode: bad shape");
  ctc_greedy_kernel<<<B, 256, 0, stream>>>(logits, seq_len, T, B, C, blank, out_labels, out_len);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_softmax_rows(const float* x, float* y, int64_t rows, int C,
                               b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(x && y && rows > 0 && C > 0, "b2_softmax_rows: bad argument");
  softmax_rows_kernel<<<cdiv(rows, 8), 256, 0, stream>>>(x, y, rows, C);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
