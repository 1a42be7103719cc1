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
                 int ldc, const float* __restrict__ bias, int k_chunk) {
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

  // split-K: slice blockIdx.z covers [kb, ke); partial sums are added atomically (C was
  // pre-scaled by beta on the host side of the launch)
  const bool split = gridDim.z > 1;
  const int kb = split ? blockIdx.z * k_chunk : 0;
  const int ke = split ? min(K, kb + k_chunk) : K;
  for (int k0 = kb; k0 < ke; k0 += BK) {
    // A tile: BM x BK  (2048 elements, 8 per thread)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = tid + r * 256;
      int m, k;
      if (TA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < ke) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
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
      if (gn < N && gk < ke) v = TB ? Bm[(int64_t)gn * ldb + gk] : Bm[(int64_t)gk * ldb + gn];
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
      float* c = C + (int64_t)gm * ldc + gn;
      if (split) {
        if (bias && blockIdx.z == 0) v += bias[gn];
        atomicAdd(c, v);
      } else {
        if (bias) v += bias[gn];
        if (beta != 0.f) v += beta * (*c);
        *c = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
scale_matrix_kernel(float* __restrict__ C, int M, int N, int ldc, float beta) {
  const int64_t n = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float* c = C + (i / N) * ldc + (i % N);
    *c = beta == 0.f ? 0.f : beta * (*c);
  }
}

int num_sms();

// ---------------------------------------------------------------------------------------
// Skinny GEMM: M <= 64 rows (one decoder step's batch).  The 128x64 tile kernel above runs such a
// product on a handful of CTAs with a serial K loop (measured 80 us for [8x1344].[1344x1024]);
// here the work is split over N tiles x K slices so that ~2 waves of CTAs stream B once
// (L2/HBM-bound, a few microseconds).  Tile: 64 columns x 64 k per iteration, B tile staged in
// shared memory as Bs[k][n] whatever its storage order, thread = (column, quarter of the rows).
// Partial sums of the K slices are added atomically (C pre-scaled by beta and seeded with the bias).
// ---------------------------------------------------------------------------------------
constexpr int SN = 64, SK = 64, SM_MAX = 64;

template <bool TB, int RPT>      // RPT rows per thread: the kernel covers M <= 4*RPT
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                   const float* __restrict__ Bm, int ldb, float beta, float* __restrict__ C, int ldc,
                   const float* __restrict__ bias, int k_chunk, int64_t strideA, int64_t strideB,
                   int64_t strideC) {
  A += blockIdx.z * strideA; Bm += blockIdx.z * strideB; C += blockIdx.z * strideC;   // batched form
  __shared__ __align__(16) float As[SM_MAX][SK + 4];   // row stride 68 floats: float4 reads along k
  __shared__ float Bs[SK][SN + 1];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * SN;
  const int kb = blockIdx.y * k_chunk, ke = min(K, kb + k_chunk);
  const int nl = tid & 63, mg = tid >> 6;              // column within the tile, row group (rows mg, mg+4, ...)
  float acc[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) acc[i] = 0.f;
  for (int k0 = kb; k0 < ke; k0 += SK) {
    // A chunk: M x SK
    for (int e = tid; e < M * SK; e += 256) {
      const int m = e / SK, k = e % SK;
      As[m][k] = (k0 + k < ke) ? A[(int64_t)m * lda + k0 + k] : 0.f;
    }
    // B tile: SK x SN
#pragma unroll
    for (int r = 0; r < (SK * SN) / 256; ++r) {
      const int e = tid + r * 256;
      int k, n;
      if (TB) { k = e % SK; n = e / SK; } else { n = e % SN; k = e / SN; }
      float v = 0.f;
      if (n0 + n < N && k0 + k < ke) v = TB ? Bm[(int64_t)(n0 + n) * ldb + k0 + k] : Bm[(int64_t)(k0 + k) * ldb + n0 + n];
      Bs[k][n] = v;
    }
    __syncthreads();
    // four k per pass: one 16-byte broadcast read of A per row instead of four scalar ones (the loop was
    // bound by shared-memory issue: 17 loads per 16 FMAs at M = 64)
#pragma unroll 4
    for (int k = 0; k < SK; k += 4) {
      const float b0 = Bs[k][nl], b1 = Bs[k + 1][nl], b2 = Bs[k + 2][nl], b3 = Bs[k + 3][nl];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float4 a4 = *(const float4*)&As[mg + 4 * i][k];
        acc[i] = fmaf(a4.x, b0, acc[i]); acc[i] = fmaf(a4.y, b1, acc[i]);
        acc[i] = fmaf(a4.z, b2, acc[i]); acc[i] = fmaf(a4.w, b3, acc[i]);
      }
    }
    __syncthreads();
  }
  const int gn = n0 + nl;
  if (gn >= N) return;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int m = mg + 4 * i;
    if (m >= M) continue;
    float v = alpha * acc[i];
    float* c = C + (int64_t)m * ldc + gn;
    if (gridDim.y > 1) atomicAdd(c, v);
    else {
      if (bias) v += bias[gn];
      if (beta != 0.f) v += beta * (*c);
      *c = v;
    }
  }
}

// C = beta*C + bias (the seed the split-K partial sums are added to)
__global__ void __launch_bounds__(256)
seed_matrix_kernel(float* __restrict__ C, int M, int N, int ldc, float beta, const float* __restrict__ bias) {
  const int64_t n = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % N);
    float* c = C + (i / N) * ldc + col;
    float v = beta == 0.f ? 0.f : beta * (*c);
    if (bias) v += bias[col];
    *c = v;
  }
}

static int gemm_skinny(int transb, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                       int ldb, float beta, float* C, int ldc, const float* bias, cudaStream_t stream) {
  const int ntiles = cdiv(N, SN);
  int splits = cdiv(2 * num_sms(), ntiles);
  if (splits > cdiv(K, SK)) splits = cdiv(K, SK);
  if (splits < 1) splits = 1;
  const int k_chunk = cdiv(cdiv(K, splits), SK) * SK;
  dim3 grid(ntiles, cdiv(K, k_chunk));
  if (grid.y > 1) {
    int blocks = cdiv((int64_t)M * N, 256); if (blocks > 1184) blocks = 1184;
    seed_matrix_kernel<<<blocks, 256, 0, stream>>>(C, M, N, ldc, beta, bias);
    B2_LAUNCH_CHECK();
  }
#define B2_SKINNY(TBv, RPTv) gemm_skinny_kernel<TBv, RPTv><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_chunk, 0, 0, 0)
  if (transb) { if (M <= 16) B2_SKINNY(true, 4); else if (M <= 32) B2_SKINNY(true, 8); else B2_SKINNY(true, 16); }
  else { if (M <= 16) B2_SKINNY(false, 4); else if (M <= 32) B2_SKINNY(false, 8); else B2_SKINNY(false, 16); }
#undef B2_SKINNY
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// batch of independent skinny products C_i = A_i . B_i (M <= 32), one K sweep per CTA (no split-K)
int gemm_skinny_batched(int M, int N, int K, const float* A, int lda, int64_t strideA, const float* B, int ldb,
                        int64_t strideB, float* C, int ldc, int64_t strideC, int batch, cudaStream_t stream) {
  if (M > SM_MAX) { set_error("gemm_skinny_batched: M=%d > %d", M, SM_MAX); return B2_ERR_INVALID; }
  dim3 grid(cdiv(N, SN), 1, batch);
  if (M <= 16)
    gemm_skinny_kernel<false, 4><<<grid, 256, 0, stream>>>(M, N, K, 1.f, A, lda, B, ldb, 0.f, C, ldc, nullptr,
                                                          cdiv(K, SK) * SK, strideA, strideB, strideC);
  else if (M <= 32)
    gemm_skinny_kernel<false, 8><<<grid, 256, 0, stream>>>(M, N, K, 1.f, A, lda, B, ldb, 0.f, C, ldc, nullptr,
                                                          cdiv(K, SK) * SK, strideA, strideB, strideC);
  else
    gemm_skinny_kernel<false, 16><<<grid, 256, 0, stream>>>(M, N, K, 1.f, A, lda, B, ldb, 0.f, C, ldc, nullptr,
                                                           cdiv(K, SK) * SK, strideA, strideB, strideC);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// two independent skinny products of the same shape in ONE launch (forward / backward direction of a
// recurrence step): C_i = A_i . op(B_i), i = 0,1; C_1 must follow C_0 (C_1 == C_0 + M*ldc).  Split-K as above.
int gemm_skinny_pair(int transb, int M, int N, int K, const float* A0, const float* A1, int lda,
                     const float* B0, const float* B1, int ldb, float* C0, float* C1, int ldc,
                     cudaStream_t stream) {
  if (M > SM_MAX || C1 != C0 + (int64_t)M * ldc) { set_error("gemm_skinny_pair: bad arguments"); return B2_ERR_INVALID; }
  const int ntiles = cdiv(N, SN);
  int splits = cdiv(num_sms(), ntiles);            // two problems share the machine
  if (splits > cdiv(K, SK)) splits = cdiv(K, SK);
  if (splits < 1) splits = 1;
  const int k_chunk = cdiv(cdiv(K, splits), SK) * SK;
  dim3 grid(ntiles, cdiv(K, k_chunk), 2);
  if (grid.y > 1) B2_CUDA(cudaMemsetAsync(C0, 0, (size_t)2 * M * ldc * sizeof(float), stream));
  const int64_t sA = A1 - A0, sB = B1 - B0, sC = C1 - C0;
#define B2_SKINNY2(TBv, RPTv) gemm_skinny_kernel<TBv, RPTv><<<grid, 256, 0, stream>>>(M, N, K, 1.f, A0, lda, B0, ldb, 0.f, C0, ldc, nullptr, k_chunk, sA, sB, sC)
  if (transb) { if (M <= 16) B2_SKINNY2(true, 4); else if (M <= 32) B2_SKINNY2(true, 8); else B2_SKINNY2(true, 16); }
  else { if (M <= 16) B2_SKINNY2(false, 4); else if (M <= 32) B2_SKINNY2(false, 8); else B2_SKINNY2(false, 16); }
#undef B2_SKINNY2
  B2_LAUNCH_CHECK();
  return B2_OK;
}

int gemm_simt(int transa, int transb, int M, int N, int K, float alpha, const float* A, int lda,
              const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
              cudaStream_t stream) {
  if (!transa && M <= SM_MAX)
    return gemm_skinny(transb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, stream);
  dim3 grid(cdiv(N, BN), cdiv(M, BM));
  // few output tiles and a long reduction (weight gradients over millions of rows): split K
  int k_chunk = K;
  const int tiles = grid.x * grid.y;
  if (tiles * 2 <= num_sms() && K >= 8192) {
    int splits = cdiv(2 * num_sms(), tiles);
    if (splits > cdiv(K, 2048)) splits = cdiv(K, 2048);
    if (splits > 1) {
      k_chunk = cdiv(cdiv(K, splits), BK) * BK;
      grid.z = cdiv(K, k_chunk);
      if (grid.z > 1 && beta != 1.f) {
        int blocks = cdiv((int64_t)M * N, 256); if (blocks > 1184) blocks = 1184;
        scale_matrix_kernel<<<blocks, 256, 0, stream>>>(C, M, N, ldc, beta);
        B2_LAUNCH_CHECK();
      }
    }
  }
  if (!transa && !transb)
    gemm_simt_kernel<false, false><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_chunk);
  else if (!transa && transb)
    gemm_simt_kernel<false, true><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_chunk);
  else if (transa && !transb)
    gemm_simt_kernel<true, false><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_chunk);
  else
    gemm_simt_kernel<true, true><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_chunk);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

}  // namespace b2
