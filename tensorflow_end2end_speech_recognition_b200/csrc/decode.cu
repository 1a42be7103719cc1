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

// ---------------------------------------------------------------------------------------
// Label error rate: Levenshtein distance of B (hypothesis, reference) pairs, one warp per pair.
// Replaces tf.edit_distance(hyp, truth, normalize=True) as used by compute_ler
// (models/ctc/ctc.py:382-398, attention_seq2seq.py:701-724).  Integer work, bit-exact.
// Row DP over the hypothesis; the reference row lives in shared memory (two int rows), lanes
// sweep the row in chunks of 32 with the left-neighbour dependency resolved by a warp scan:
//   d[j] = min(up[j] + 1, diag[j] + cost, d[j-1] + 1)  ->  e[j] = min(up+1, diag+cost) then
//   d[j] = min_k<=j (e[k] + (j - k))  (prefix-min of e[k]-k, plus j).
// ---------------------------------------------------------------------------------------
namespace b2 {

__global__ void __launch_bounds__(128)
edit_distance_kernel(const int* __restrict__ hyp, const int* __restrict__ hyp_off,
                     const int* __restrict__ ref, const int* __restrict__ ref_off, int B, int max_ref,
                     int* __restrict__ dist) {
  extern __shared__ int sm_ed[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * 4 + warp;
  if (p >= B) return;
  int* row0 = sm_ed + (size_t)warp * 2 * (max_ref + 1);
  int* row1 = row0 + (max_ref + 1);
  const int* h = hyp + hyp_off[p];
  const int* r = ref + ref_off[p];
  const int n = hyp_off[p + 1] - hyp_off[p], m = ref_off[p + 1] - ref_off[p];
  for (int j = lane; j <= m; j += 32) row0[j] = j;
  __syncwarp();
  int* up = row0; int* cur = row1;
  for (int i = 1; i <= n; ++i) {
    const int hc = h[i - 1];
    int carry = i;                                   // d[i][0]
    if (lane == 0) cur[0] = i;
    for (int j0 = 1; j0 <= m; j0 += 32) {
      const int j = j0 + lane;
      int e = 0x3fffffff;
      if (j <= m) e = min(up[j] + 1, up[j - 1] + (hc != r[j - 1] ? 1 : 0));
      // prefix-min of (e[k] - k) over the chunk, seeded with the carry from the previous chunk
      int v = e - j;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = min(v, t);
      }
      int d = min(v + j, carry + (j - (j0 - 1)));
      if (j <= m) cur[j] = d;
      carry = __shfl_sync(0xffffffffu, d, 31);
    }
    __syncwarp();
    int* t = up; up = cur; cur = t;
  }
  if (lane == 0) dist[p] = up[m];
}

}  // namespace b2

extern "C" int b2_edit_distance(const int32_t* hyp, const int32_t* hyp_offsets, const int32_t* ref,
                                const int32_t* ref_offsets, int B, int max_ref_len, int32_t* dist,
                                b2_stream_t stream_) {
  using namespace b2;
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(hyp_offsets && ref_offsets && dist && B > 0 && max_ref_len >= 0, "b2_edit_distance: bad argument");
  const size_t smem = (size_t)4 * 2 * (max_ref_len + 1) * sizeof(int);
  B2_CHECK_ARG(smem <= 200 * 1024, "b2_edit_distance: reference of %d labels too long", max_ref_len);
  B2_CUDA(cudaFuncSetAttribute(edit_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  edit_distance_kernel<<<cdiv(B, 4), 128, smem, stream>>>(hyp, hyp_offsets, ref, ref_offsets, B, max_ref_len, dist);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
