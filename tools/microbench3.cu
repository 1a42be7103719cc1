// Probe 4: cost structure of small-N tcgen05.mma chains (issue vs dependency latency)
// Probe 5: issue cost of cp.async.bulk smem->cluster-smem copies from 1 warp vs 4 warps
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../tensorflow_end2end_speech_recognition_b200/csrc/sm100.cuh"
using namespace b2::sm100;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// one CTA: KS chained MMAs of shape 128 x N x 16 per "step", NACC accumulators, TS or SS
__global__ void __launch_bounds__(128, 1)
k_mma_chain(int N, int KS, int NACC, int ts_mode, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                       // no-swizzle K-major [kc][ng][8][16B], up to N=256, K=16*KS
  uint8_t* sA = smem + 65536;               // no-swizzle K-major A for SS mode: 128 x 16 per k-step (re-used)
  uint64_t* bar = (uint64_t*)(smem + 65536 + 4096 * 32);
  uint32_t* slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 4096 * 32) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t ng = N / 8;
    long long t_issue = 0, t_total = 0;
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int k = 0; k < KS; ++k) {
        // B: per k-step 2 kc chunks, each ng*128 bytes
        const uint64_t bd = make_smem_desc(smem_u32(sB) + (k % 8) * 2 * ng * 128, ng * 128, 128, 0);
        const uint32_t acc = tb + 256 + (k % NACC) * N;     // N <= 64 when NACC = 4
        if (ts_mode) mma_ts(acc, tb + (k % 32) * 8, bd, idesc, k >= NACC ? 1u : 0u);
        else {
          const uint64_t ad = make_smem_desc(smem_u32(sA) + (k % 32) * 4096, 2048, 128, 0);
          mma_ss(acc, ad, bd, idesc, k >= NACC ? 1u : 0u);
        }
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

// cluster of 16: time to issue + complete 16 bulk copies of `slice` bytes with different issuers
template <int PATTERN>   // 0: 16 lanes of warp 0 ; 1: 4 lanes in each of 4 warps ; 2: 1 lane in each of 16 warps... (8 warps here: 2 each)
__global__ void __launch_bounds__(256, 1)
k_bulk_issue(int iters, int slice, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = cluster_ctarank();
  const uint32_t csize = 16;
  uint8_t* buf = smem;
  uint8_t* stage = smem + 2 * csize * slice;
  uint64_t* full = (uint64_t*)(stage + slice);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_mbar_init(); }
  for (int i = tid; i < slice / 4; i += blockDim.x) ((uint32_t*)stage)[i] = rank;
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync();
  uint32_t ph[2] = {0, 0};
  long long t_issue = 0, t_all = 0;
  for (int it = 0; it < iters; ++it) {
    const int p = it & 1;
    uint8_t* dst = buf + (size_t)p * csize * slice + rank * slice;
    const long long t0 = clock64();
    if (tid == 0) mbar_expect_tx(&full[p], csize * slice);
    int d = -1;
    if (PATTERN == 0) { if (warp == 0 && lane < 16) d = lane; }
    else if (PATTERN == 1) { if (warp < 4 && lane < 4) d = warp * 4 + lane; }
    else { if (lane < 2) d = warp * 2 + lane; }
    if (d >= 0) bulk_s2cluster(dst, stage, slice, &full[p], (uint32_t)d);
    const long long t1 = clock64();
    mbar_wait_cluster(&full[p], ph[p]); ph[p] ^= 1;
    const long long t2 = clock64();
    if (tid == 0 && it > 2) { t_issue += t1 - t0; t_all += t2 - t0; }
    __syncthreads();
  }
  cluster_sync();
  if (tid == 0 && rank == 0) { out[0] = t_issue / (iters - 3); out[1] = t_all / (iters - 3); }
}

template <int PATTERN>
static void run_bulk(int slice) {
  long long* d_out; CK(cudaMalloc(&d_out, 16));
  size_t smem = 2 * 16 * (size_t)slice + slice + 64;
  CK(cudaFuncSetAttribute(k_bulk_issue<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_bulk_issue<PATTERN>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(16); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k_bulk_issue<PATTERN>, 500, slice, d_out));
  CK(cudaDeviceSynchronize());
  long long h[2]; CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
  printf("[bulk issue pattern=%d slice=%d] issue %lld cycles, issue+complete %lld cycles\n", PATTERN, slice, h[0], h[1]);
  CK(cudaFree(d_out));
}

int main() {
  long long* d_out; CK(cudaMalloc(&d_out, 16));
  const size_t smem = 65536 + 4096 * 32 + 64;
  CK(cudaFuncSetAttribute(k_mma_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int ts = 1; ts >= 0; --ts)
    for (int N : {16, 32, 64, 128, 256})
      for (int nacc : {1, 2, 4}) {
        if (nacc * N > 256) continue;
        k_mma_chain<<<1, 128, smem>>>(N, 32, nacc, ts, 50, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma chain failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2]; CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("[mma chain %s N=%3d KS=32 NACC=%d] issue %5lld cycles  issue+complete %5lld cycles  (%.1f per MMA)\n",
               ts ? "TS" : "SS", N, nacc, h[0], h[1], h[1] / 32.0);
      }
  for (int slice : {512, 1024, 2048}) { run_bulk<0>(slice); run_bulk<1>(slice); run_bulk<2>(slice); }
  return 0;
}
