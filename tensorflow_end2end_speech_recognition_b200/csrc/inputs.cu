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

__global__ void __launch_bounds__(256)
stack_splice_kernel(const float* __restrict__ raw, const int* __restrict__ raw_len, int B, int Traw, int D,
                    int S, int K, int P, int Tout, int Dout, float* __restrict__ out,
                    int* __restrict__ out_len) {
  const int64_t total = (int64_t)B * Tout * Dout;
  const int R = P * S;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % Dout);
    const int j = (int)((idx / Dout) % Tout);
    const int b = (int)(idx / ((int64_t)Dout * Tout));
    const int len = min(raw_len[b], Traw);
    const int olen = (S == 1) ? len : (len + K - 1) / K;
    if (o == 0 && j == 0) out_len[b] = min(olen, Tout);
    float v = 0.f;
    if (j < olen) {
      int js = j, e = o;
      bool ok = true;
      if (P > 1) {
        const int c = o / (R * 3), r = (o / 3) % R, k = o % 3;
        if (r <= P - 1) { js = max(0, j + r - P); e = (c * 3 + k) * S; }
        else if (r <= P - 2 + S) { js = max(0, j - 1); e = (c * 3 + k) * S + (r - P + 1); }
        else ok = false;
      }
      if (ok) {
        int t = js, d = e;
        if (S > 1) { t = js * K + e / D; d = e % D; }
        if (t < len) v = raw[((size_t)b * Traw + t) * D + d];
      }
    }
    out[idx] = v;
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
  const int64_t total = (int64_t)B * Tout * Dout;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  stack_splice_kernel<<<(int)blocks, 256, 0, stream>>>(raw, raw_len, B, Traw, D, num_stack, num_skip, splice,
                                                      Tout, Dout, out, out_len);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
