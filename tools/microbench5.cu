// Probe 7: does issuing tcgen05.mma from several warps overlap the ~53-cycle per-instruction floor?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../tensorflow_end2end_speech_recognition_b200/csrc/sm100.cuh"
using namespace b2::sm100;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int N, int NW, int MODE>
__global__ void __launch_bounds__(128, 1)
k_mma_mw(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sA = smem + 65536;
  uint64_t* bar = (uint64_t*)(smem + 65536 + 131072);   // [4]
  uint32_t* slot = (uint32_t*)(bar + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (65536 + 131072) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  constexpr uint32_t ng = N / 8;
  const uint64_t bd0 = make_smem_desc(smem_u32(sB), ng * 128, 128, 0);
  const uint64_t ad0 = make_smem_desc(smem_u32(sA), 16 * 128, 128, 0);
  long long t_total = 0;
  uint32_t ph = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    if (warp < NW && lane == 0) {
      constexpr int PER = 32 / NW;
#pragma unroll
      for (int kk = 0; kk < PER; ++kk) {
        const int k = warp * PER + kk;
        if (MODE == 0) mma_ts(tb + 256 + warp * 64, tb + k * 8, bd0 + (uint64_t)((k % 8) * (2 * ng * 128 / 16)), idesc, kk > 0);
        else mma_ss(tb + 256 + warp * 64, ad0 + (uint64_t)(k * (2 * 16 * 128 / 16)), bd0 + (uint64_t)((k % 8) * (2 * ng * 128 / 16)), idesc, kk > 0);
      }
      mma_commit(&bar[warp]);
    }
    if (tid == 0) {
      for (int w = 0; w < NW; ++w) mbar_wait(&bar[w], ph);
      const long long t2 = clock64();
      if (r > 0) t_total += t2 - t0;
    }
    ph ^= 1;
  }
  if (tid == 0) out[0] = t_total / (reps - 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int NW, int MODE>
static void run(long long* d_out) {
  const size_t smem = 65536 + 131072 + 128;
  CK(cudaFuncSetAttribute(k_mma_mw<N, NW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mma_mw<N, NW, MODE><<<1, 128, smem>>>(50, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d NW=%d failed: %s\n", N, NW, cudaGetErrorString(e)); exit(1); }
  long long h[1]; CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
  printf("[mma multi-warp %s N=%3d issuing warps=%d] 32 MMAs complete in %5lld cycles\n", MODE ? "SS" : "TS", N, NW, h[0]);
}

int main() {
  long long* d_out; CK(cudaMalloc(&d_out, 16));
  run<16, 1, 0>(d_out); run<16, 2, 0>(d_out); run<16, 4, 0>(d_out);
  run<32, 1, 0>(d_out); run<32, 2, 0>(d_out); run<32, 4, 0>(d_out);
  run<16, 1, 1>(d_out); run<16, 2, 1>(d_out); run<16, 4, 1>(d_out);
  return 0;
}
