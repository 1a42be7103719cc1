// fp32 CUDA-core GEMM (B2_PREC_FP32): the deterministic parity path behind
// b2_gemm.  C = alpha*op(A).op(B) + beta*C + bias, any M/N/K, any transposes.
// 128x64 CTA tile, BK=16, 256 threads, 8x4 register micro-tile, smem staged.
// The throughput path is gemm_tcgen05.cu; this one exists so that every
// GEMM-shaped op has an fp32-exact twin to pin parity against the oracle.
#include "common.cuh"

namespace b2 {

constexpr int BM = 128, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                 const float* __restrict__ Bm, int ldb, float beta, float* __restrict__ C,
                 int ldc, const float* __restrict__ bias) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads; thread tile 8 (m) x 4 (n)
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: BM x BK  (2048 elements, 8 per thread)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = tid + r * 256;
      int m, k;
      if (TA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
      As[k][m] = v;
    }
    // B tile: BK x BN  (1024 elements, 4 per thread)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + r * 256;
      int n, k;
      if (TB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = TB ? Bm[(int64_t)gn * ldb + gk] : Bm[(int64_t)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[k][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[gn];
      float* c = C + (int64_t)gm * ldc + gn;
      if (beta != 0.f) v += beta * (*c);
      *c = v;
    }
  }
}

int gemm_simt(int transa, int transb, int M, int N, int K, float alpha, const float* A, int lda,
              const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
              cudaStream_t stream) {
  dim3 grid(cdiv(N, BN), cdiv(M, BM));
  if (!transa && !transb)
    gemm_simt_kernel<false, false><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else if (!transa && transb)
    gemm_simt_kernel<false, true><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else if (transa && !transb)
    gemm_simt_kernel<true, false><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else
    gemm_simt_kernel<true, true><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

}  // namespace b2
