// Hardware probes that decide the design of the persistent BLSTM recurrence
// kernel (run on the GPU box via gpurun; results recorded in DESIGN.md):
//   1. can a 16-CTA cluster be co-scheduled 8x on this part?
//   2. cost of one all-gather step of h over DSMEM (bulk copy vs st.shared::cluster)
//   3. tcgen05.mma with A in TMEM: operand packing + no-swizzle K-major B descriptor
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../tensorflow_end2end_speech_recognition_b200/csrc/sm100.cuh"
using namespace b2::sm100;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- probe 2
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gsrc, uint32_t bytes,
                                                   uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
      "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// 0: bulk copy smem->cluster smem + complete_tx ; 1: st.shared::cluster + arrive ;
// 2: st.global own slice + fence.proxy.async + multicast bulk load global->all CTAs
template <int MODE>
__global__ void __launch_bounds__(192, 1)
k_allgather(int iters, int slice, long long* out, uint8_t* gbuf) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = cluster_ctarank();
  uint32_t csize;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  uint8_t* buf = smem;                          // [2][csize*slice]
  uint8_t* stage = smem + 2 * csize * slice;    // [slice]
  uint64_t* full = (uint64_t*)(stage + slice);  // [2]
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&full[0], MODE == 1 ? csize : 1);
    mbar_init(&full[1], MODE == 1 ? csize : 1);
    fence_mbar_init();
  }
  for (int i = tid; i < slice / 4; i += blockDim.x) ((uint32_t*)stage)[i] = rank * 1000 + i;
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  uint32_t ph[2] = {0, 0};
  for (int it = 0; it < iters; ++it) {
    const int p = it & 1;
    uint8_t* dst = buf + (size_t)p * csize * slice + rank * slice;
    if (MODE == 0) {
      if (tid == 0) mbar_expect_tx(&full[p], csize * slice);
      if (tid < csize) bulk_s2cluster(dst, stage, slice, &full[p], tid);
    } else if (MODE == 2) {
      uint8_t* g = gbuf + ((size_t)(blockIdx.x / csize) * 2 + p) * csize * slice + rank * slice;
      if (tid == 0) mbar_expect_tx(&full[p], csize * slice);
      if (tid < 64) {
        for (int o = tid * 16; o < slice; o += 64 * 16) {
          uint4 v = *(const uint4*)(stage + o);
          v.x += it;
          *(uint4*)(g + o) = v;
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("bar.sync 1, 64;" ::: "memory");
        if (tid == 0) bulk_g2s_multicast(dst, g, slice, &full[p], (uint16_t)((1u << csize) - 1));
      }
    } else {
      // 64 threads x 16 B = 1 KB per pass over the slice
      const uint32_t d0 = smem_u32(dst);
      for (uint32_t c = 0; c < csize; ++c) {
        const uint32_t base = mapa(d0, c);
        for (int o = tid * 16; o < slice; o += 64 * 16)
          if (tid < 64) {
            const uint4 v = *(const uint4*)(stage + o);
            st_cluster_v4(base + o, v.x, v.y, v.z, v.w);
          }
      }
      // make the stores visible, then one arrive per destination CTA
      if (tid < 64) {
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        asm volatile("bar.sync 1, 64;" ::: "memory");
        if (tid < csize) mbar_arrive_cluster(&full[p], tid);
      }
    }
    if (tid >= 64 || MODE == 0) { /* all threads wait */ }
    mbar_wait_cluster(&full[p], ph[p]);
    ph[p] ^= 1;
    if (MODE != 0) asm volatile("bar.sync 2, 192;" ::: "memory");
  }
  long long t1 = clock64();
  // checksum so nothing is optimised away
  uint32_t s = 0;
  for (int i = tid; i < (int)(csize * slice / 4); i += blockDim.x) s += ((uint32_t*)buf)[i];
  cluster_sync();
  if (tid == 0 && rank == 0) { out[blockIdx.x / csize * 2] = t1 - t0; out[blockIdx.x / csize * 2 + 1] = s; }
}

template <int MODE>
static void run_allgather(int csize, int nclusters, int slice, int iters) {
  long long* d_out; CK(cudaMalloc(&d_out, 64 * sizeof(long long)));
  uint8_t* gbuf; CK(cudaMalloc(&gbuf, (size_t)64 * 2 * 16 * 8192)); CK(cudaMemset(gbuf, 0, (size_t)64 * 2 * 16 * 8192));
  CK(cudaMemset(d_out, 0, 64 * sizeof(long long)));
  size_t smem = 2 * (size_t)csize * slice + slice + 64;
  CK(cudaFuncSetAttribute(k_allgather<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_allgather<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * nclusters); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int maxc = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, k_allgather<MODE>, &cfg);
  printf("[allgather mode=%d csize=%d slice=%d smem=%zu] maxActiveClusters=%d (%s)\n", MODE, csize, slice, smem, maxc, cudaGetErrorString(e));
  if (maxc < nclusters) { printf("  -> cannot co-schedule %d clusters, skipping\n", nclusters); cudaGetLastError(); return; }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    e = cudaLaunchKernelEx(&cfg, k_allgather<MODE>, iters, slice, d_out, gbuf);
    if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return; }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[64]; CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("  %d clusters x %d iters: %.3f ms total, %.1f ns/iter, cluster0 %.0f cycles/iter (checksum %lld)\n",
         nclusters, iters, ms, ms * 1e6 / iters, (double)h[0] / iters, h[1]);
  CK(cudaFree(d_out)); CK(cudaFree(gbuf));
}

// ---------------------------------------------------------------- probe 3
// D[128 x 16] = A[128 x 16] . B[16 x 16]^T-ish : A in TMEM (TS) and in smem (SS) for comparison
__global__ void __launch_bounds__(128, 1)
k_ts_mma(const __nv_bfloat16* A /*[128][16]*/, const __nv_bfloat16* Bm /*[16 n][16 k]*/, float* D_ts /*[2][128][16]*/, float* D_ss, int swap_lbo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                 // no-swizzle K-major: [kc(2)][ng(2)][8 rows][16 B]
  uint8_t* sA = smem + 1024;          // no-swizzle K-major: [kc(2)][mg(16)][8 rows][16 B]
  uint64_t* bar = (uint64_t*)(smem + 1024 + 4096);
  uint32_t* slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // stage B and A into the canonical no-swizzle layout
  for (int e = tid; e < 16 * 16; e += 128) {
    const int n = e / 16, k = e % 16;
    const int off = (k / 8) * 256 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
    *(__nv_bfloat16*)(sB + off) = Bm[e];
  }
  for (int e = tid; e < 128 * 16; e += 128) {
    const int m = e / 16, k = e % 16;
    const int off = (k / 8) * 2048 + (m / 8) * 128 + (m % 8) * 16 + (k % 8) * 2;
    *(__nv_bfloat16*)(sA + off) = A[e];
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  const uint32_t tA = tb + 32;       // columns 32.. hold A (8 columns), D at columns 0..15
  const uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
  for (int variant = 0; variant < 2; ++variant) {
    // variant 0: low half = even k ; variant 1: low half = odd k
    uint32_t r[8];
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint16_t lo = __bfloat16_as_ushort(A[row * 16 + 2 * j + (variant ? 1 : 0)]);
      const uint16_t hi = __bfloat16_as_ushort(A[row * 16 + 2 * j + (variant ? 0 : 1)]);
      r[j] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
    tmem_st_32x32b_x8(tA + ((uint32_t)(warp * 32) << 16), r);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      const uint64_t bd = swap_lbo ? make_smem_desc(smem_u32(sB), 128, 256, 0) : make_smem_desc(smem_u32(sB), 256, 128, 0);
      mma_ts(tb, tA, bd, idesc, 0);
      mma_commit(bar);
    }
    mbar_wait(bar, variant & 1);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld_32x32b_x16(tb + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D_ts[(variant * 128 + row) * 16 + j] = __uint_as_float(v[j]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // SS reference run
  if (tid == 0) {
    const uint64_t bd = swap_lbo ? make_smem_desc(smem_u32(sB), 128, 256, 0) : make_smem_desc(smem_u32(sB), 256, 128, 0);
    const uint64_t ad = swap_lbo ? make_smem_desc(smem_u32(sA), 128, 2048, 0) : make_smem_desc(smem_u32(sA), 2048, 128, 0);
    mma_ss(tb, ad, bd, idesc, 0);
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  {
    uint32_t v[16];
    const int row = warp * 32 + lane;
    tmem_ld_32x32b_x16(tb + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D_ss[row * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static void run_ts() {
  std::vector<__nv_bfloat16> A(128 * 16), B(16 * 16);
  std::vector<float> Af(128 * 16), Bf(16 * 16);
  srand(1);
  for (int i = 0; i < 128 * 16; ++i) { Af[i] = bf((rand() % 2001 - 1000) / 500.f); A[i] = __float2bfloat16(Af[i]); }
  for (int i = 0; i < 16 * 16; ++i) { Bf[i] = bf((rand() % 2001 - 1000) / 500.f); B[i] = __float2bfloat16(Bf[i]); }
  std::vector<float> ref(128 * 16);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) { float s = 0; for (int k = 0; k < 16; ++k) s += Af[m * 16 + k] * Bf[n * 16 + k]; ref[m * 16 + n] = s; }
  __nv_bfloat16 *dA, *dB; float *dts, *dss;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dts, 2 * 128 * 16 * 4)); CK(cudaMalloc(&dss, 128 * 16 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  for (int swap = 0; swap < 1; ++swap) {   // swap=1 (LBO<->SBO exchanged) faults: confirmed wrong on B200
    CK(cudaMemset(dts, 0, 2 * 128 * 16 * 4)); CK(cudaMemset(dss, 0, 128 * 16 * 4));
    k_ts_mma<<<1, 128, 8192>>>(dA, dB, dts, dss, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[ts_mma swap=%d] kernel failed: %s\n", swap, cudaGetErrorString(e)); return; }
    std::vector<float> ts(2 * 128 * 16), ss(128 * 16);
    CK(cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost));
    double e0 = 0, e1 = 0, es = 0;
    for (int i = 0; i < 128 * 16; ++i) { e0 = fmax(e0, fabs(ts[i] - ref[i])); e1 = fmax(e1, fabs(ts[128 * 16 + i] - ref[i])); es = fmax(es, fabs(ss[i] - ref[i])); }
    printf("[ts_mma swap_lbo_sbo=%d] max|err|: TS(lo=even k)=%.4g  TS(lo=odd k)=%.4g  SS=%.4g\n", swap, e0, e1, es);
  }
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs=%d smemOptin=%zu\n", p.name, p.major, p.minor, p.multiProcessorCount, p.sharedMemPerBlockOptin);
  run_ts();
  const int iters = 2000;
  for (int cs : {16, 8}) {
    for (int slice : {1024, 2048}) {
      run_allgather<0>(cs, 1, slice, iters);
      run_allgather<2>(cs, 1, slice, iters);
      run_allgather<0>(cs, 4, slice, iters);
      run_allgather<2>(cs, 4, slice, iters);
      run_allgather<2>(cs, cs == 16 ? 7 : 16, slice, iters);
    }
  }
  return 0;
}
