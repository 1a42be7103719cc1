// Probe 8: does a thread that is issuing tcgen05.mma back-to-back slow down OTHER warps of
// the same SM sub-partition (issue-port blocking)?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../tensorflow_end2end_speech_recognition_b200/csrc/sm100.cuh"
using namespace b2::sm100;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256, 1)
k_probe(int do_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = (uint64_t*)(smem + 65536);
  uint32_t* slot = (uint32_t*)(bar + 1);
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += 256) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (warp == 0) {
    if (lane == 0 && do_mma) {
      const uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
      const uint64_t bd0 = make_smem_desc(smem_u32(smem), 256, 128, 0);
      uint32_t ph = 0;
      while (!stop) {
#pragma unroll
        for (int k = 0; k < 32; ++k) mma_ts(tb + 256, tb + k * 8, bd0 + (uint64_t)((k % 8) * 32), idesc, k > 0);
        mma_commit(bar);
        mbar_wait(bar, ph); ph ^= 1;
      }
    }
  } else if (warp == 4 || warp == 5) {
    // warp 4 shares the sub-partition of warp 0 (wid % 4 == 0); warp 5 does not
    float a = lane * 0.001f, b = 1.0001f, c = 0.5f, d = 0.25f;
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 2000; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) { a = fmaf(a, b, c); c = fmaf(c, b, d); d = fmaf(d, b, a); }
    }
    const long long t1 = clock64();
    if (lane == 0) { out[warp - 4] = t1 - t0; out[2 + warp - 4] = (long long)(a + c + d); }
    __syncwarp();
    if (warp == 4 && lane == 0) { __threadfence_block(); }
  }
  // let both measuring warps finish, then stop the issuer
  if (warp == 4 || warp == 5) { asm volatile("bar.sync 1, 64;"); if (tid == 128) stop = 1; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d_out; CK(cudaMalloc(&d_out, 64));
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64));
  for (int mma = 0; mma < 2; ++mma) {
    k_probe<<<1, 256, 65536 + 64>>>(mma, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[4]; CK(cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost));
    printf("[issue-port probe] issuer %s: FMA loop (96000 dependent-ish FMAs) warp4(same SMSP as issuer)=%lld cycles  warp5(other SMSP)=%lld cycles\n",
           mma ? "ISSUING MMAs" : "idle", h[0], h[1]);
  }
  return 0;
}
