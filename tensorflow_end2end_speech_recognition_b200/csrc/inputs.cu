// Input pipeline on the device: frame stacking / skipping + splicing + zero padding in ONE
// gather pass (HBM-bound byte moving; bit-exact against the reference's numpy transforms).
//
// Replaces stack_frame (utils/io/inputs/frame_stacking.py:14-85) and do_splice
// (utils/io/inputs/splicing.py:9-73) as DatasetBase.__next__ applies them per utterance before
// padding (utils/dataset/ctc.py:120-160).  Index algebra (see oracle/inputs.py for the derivation
// from the reference's loops):
//   stacked frame js, element e  = raw frame js*num_skip + e / D, feature e % D   (0 past the end)
//   spliced frame j, element o   : c = o / (R*3), r = (o / 3) % R, k = o % 3,  R = splice*num_stack
//        r <= splice-1            -> stacked frame max(0, j + r - splice), element (c*3+k)*num_stack
//        r <= splice-2+num_stack  -> stacked frame max(0, j - 1), element (c*3+k)*num_stack + r-splice+1
//        else                      -> 0
#include "common.cuh"

namespace b2 {

// The map output element o -> (frame delta, stack slot, raw feature) is the same for every output
// frame, so each CTA builds it once in shared memory (no per-element divisions afterwards) and then
// copies ROWS_PER_CTA output frames with it.  grid = ceil(B*Tout / ROWS_PER_CTA).
constexpr int kRowsPerCta = 16;

__global__ void __launch_bounds__(256)
stack_splice_kernel(const float* __restrict__ raw, const int* __restrict__ raw_len, int B, int Traw, int D,
                    int S, int K, int P, int Tout, int Dout, float* __restrict__ out,
                    int* __restrict__ out_len) {
  extern __shared__ int s_map[];        // [Dout] packed: bits 0..19 raw feature d, 20..25 stack slot, 26..30 delta+1.., 31 valid
  int* s_dj = s_map + Dout;             // [Dout] frame delta (<= 0) in stacked-frame units
  const int R = P * S;
  for (int o = threadIdx.x; o < Dout; o += 256) {
    int dj = 0, e = o, ok = 1;
    if (P > 1) {
      const int c = o / (R * 3), r = (o / 3) % R, k = o % 3;
      if (r <= P - 1) { dj = r - P; e = (c * 3 + k) * S; }
      else if (r <= P - 2 + S) { dj = -1; e = (c * 3 + k) * S + (r - P + 1); }
      else ok = 0;
    }
    const int istack = (S > 1) ? e / D : 0;
    const int d = (S > 1) ? e % D : e;
    s_map[o] = ok ? ((istack << 20) | d) : -1;
    s_dj[o] = dj;
  }
  __syncthreads();
  const int64_t rows = (int64_t)B * Tout;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerCta;
  for (int rr = 0; rr < kRowsPerCta; ++rr) {
    const int64_t row = row0 + rr;
    if (row >= rows) break;
    const int b = (int)(row / Tout), j = (int)(row % Tout);
    const int len = min(raw_len[b], Traw);
    const int olen = (S == 1) ? len : (len + K - 1) / K;
    if (j == 0 && threadIdx.x == 0) out_len[b] = min(olen, Tout);
    float* orow = out + row * Dout;
    const float* rb = raw + (size_t)b * Traw * D;
    if (j >= olen) {
      for (int o = threadIdx.x; o < Dout; o += 256) orow[o] = 0.f;
      continue;
    }
    for (int o = threadIdx.x; o < Dout; o += 256) {
      const int m = s_map[o];
      float v = 0.f;
      if (m >= 0) {
        const int js = (P > 1) ? max(0, j + s_dj[o]) : j;
        const int t = (S > 1) ? js * K + (m >> 20) : js;
        if (t < len) v = rb[(size_t)t * D + (m & 0xfffff)];
      }
      orow[o] = v;
    }
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_stack_splice_out_dim(int D, int num_stack, int splice) {
  if (splice <= 1) return D * num_stack;
  return (D / 3) * splice * num_stack * 3;
}

extern "C" int b2_stack_splice(const float* raw, const int32_t* raw_len, int B, int Traw, int D,
                               int num_stack, int num_skip, int splice, int Tout, float* out,
                               int32_t* out_len, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(raw && raw_len && out && out_len, "b2_stack_splice: null pointer");
  B2_CHECK_ARG(B > 0 && Traw > 0 && D > 0 && Tout > 0, "b2_stack_splice: bad shape");
  B2_CHECK_ARG(num_stack >= 1 && num_skip >= 1 && splice >= 1, "b2_stack_splice: bad parameters");
  B2_CHECK_ARG(num_stack == 1 || num_stack >= num_skip, "num_skip must be less than num_stack.");
  B2_CHECK_ARG(splice == 1 || D % 3 == 0, "b2_stack_splice: splicing needs a feature width divisible by 3");
  const int Dout = b2_stack_splice_out_dim(D, num_stack, splice);
  B2_CHECK_ARG(D < (1 << 20) && num_stack < 2048, "b2_stack_splice: feature width / stack too large");
  const size_t smem = (size_t)2 * Dout * sizeof(int);
  B2_CHECK_ARG(smem <= 200 * 1024, "b2_stack_splice: output frame of %d features too wide", Dout);
  B2_CUDA(cudaFuncSetAttribute(stack_splice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = ((int64_t)B * Tout + kRowsPerCta - 1) / kRowsPerCta;
  stack_splice_kernel<<<(int)blocks, 256, smem, stream>>>(raw, raw_len, B, Traw, D, num_stack, num_skip, splice,
                                                         Tout, Dout, out, out_len);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
