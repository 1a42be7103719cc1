// Probe 6: clean per-instruction cost of tcgen05.mma (fully unrolled, precomputed descriptors)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../tensorflow_end2end_speech_recognition_b200/csrc/sm100.cuh"
using namespace b2::sm100;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int M, int N, int MODE, int NACC = 1>   // MODE 0: TS (A in TMEM) ; 1: SS no-swizzle ; 2: SS 128B swizzle (K-major, 64-wide K tile)
__global__ void __launch_bounds__(128, 1)
k_mma(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                 // 64 KB region for B
  uint8_t* sA = smem + 65536;         // 128 KB region for A
  uint64_t* bar = (uint64_t*)(smem + 65536 + 131072);
  uint32_t* slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 131072) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(M, N, 0, 0);
    constexpr uint32_t ng = N / 8, mg = M / 8;
    const uint64_t bd0 = MODE == 2 ? make_smem_desc(smem_u32(sB), 16, 1024, 2)
                                   : make_smem_desc(smem_u32(sB), ng * 128, 128, 0);
    const uint64_t ad0 = MODE == 2 ? make_smem_desc(smem_u32(sA), 16, 1024, 2)
                                   : make_smem_desc(smem_u32(sA), mg * 128, 128, 0);
    long long t_issue = 0, t_total = 0;
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        constexpr int kmodB = (65536 / (2 * ng * 128)) < 32 ? (65536 / (2 * ng * 128)) : 32;
        constexpr int kmodB2 = (65536 / (N * 128)) < 8 ? (65536 / (N * 128)) : 8;
        if (MODE == 0) mma_ts(tb + 256 + (k % NACC) * 64, tb + k * 8, bd0 + (uint64_t)((k % kmodB) * (2 * ng * 128 / 16)), idesc, k >= NACC);
        else if (MODE == 1) mma_ss(tb + 256 + (k % NACC) * 64, ad0 + (uint64_t)(k * (2 * mg * 128 / 16)), bd0 + (uint64_t)((k % kmodB) * (2 * ng * 128 / 16)), idesc, k >= NACC);
        else mma_ss(tb + 256, ad0 + (uint64_t)((k / 4) * (M * 128 / 16) + (k % 4) * 2), bd0 + (uint64_t)(((k / 4) % kmodB2) * (N * 128 / 16) + (k % 4) * 2), idesc, k > 0);
      }
      mma_commit(bar);
      const long long t1 = clock64();
      mbar_wait(bar, ph); ph ^= 1;
      const long long t2 = clock64();
      if (r > 0) { t_issue += t1 - t0; t_total += t2 - t0; }
    }
    out[0] = t_issue / (reps - 1);
    out[1] = t_total / (reps - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int M, int N, int MODE, int NACC = 1>
static void run(long long* d_out) {
  const size_t smem = 65536 + 131072 + 64;
  CK(cudaFuncSetAttribute(k_mma<M, N, MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mma<M, N, MODE, NACC><<<1, 128, smem>>>(50, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M=%d N=%d mode=%d failed: %s\n", M, N, MODE, cudaGetErrorString(e)); exit(1); }
  long long h[2]; CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
  const char* nm[3] = {"TS", "SS-nosw", "SS-sw128"};
  printf("[mma M=%3d N=%3d %-8s NACC=%d x32] issue %5lld  issue+complete %5lld cycles  (%.1f per MMA)\n", M, N, nm[MODE], NACC, h[0], h[1], h[1] / 32.0);
}

int main() {
  long long* d_out; CK(cudaMalloc(&d_out, 16));
  run<128, 16, 0, 1>(d_out); run<128, 16, 0, 2>(d_out); run<128, 16, 0, 4>(d_out);
  run<128, 32, 0, 1>(d_out); run<128, 32, 0, 2>(d_out); run<128, 32, 0, 4>(d_out);
  run<128, 64, 0, 2>(d_out); run<128, 64, 0, 4>(d_out);
  run<128, 16, 1, 1>(d_out); run<128, 16, 1, 2>(d_out); run<128, 16, 1, 4>(d_out);
  run<128, 64, 1, 4>(d_out);
  return 0;
}
