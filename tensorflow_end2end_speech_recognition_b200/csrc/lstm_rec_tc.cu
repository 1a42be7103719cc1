// Persistent BLSTM recurrence kernels on tcgen05 for sm_100a (B2_PREC_BF16).
//
// The strictly sequential half of the LSTM gate GEMMs: z_t = G_t + h_{t-1} . Wh for
// T steps, and its BPTT mirror (reference: LSTMBlockCell inside
// tf.nn.bidirectional_dynamic_rnn, models/encoders/core/blstm.py:287-320).  One launch
// runs a whole layer, both directions:
//
//   * one thread-block CLUSTER of CS = H/32 CTAs per (direction, batch groups);
//     CTA `r` owns hidden units [32r, 32r+32) = 128 gate rows (unit-major, gate-minor);
//   * the CTA's slice of Wh (bf16) is loaded ONCE into TENSOR MEMORY and stays there
//     for all T steps: the step GEMM is tcgen05.mma with A from TMEM and B (16 batch
//     columns, bf16, K-major no-swizzle) from shared memory, fp32 accumulators in TMEM
//     (swap-AB: gate rows / units are the MMA M dimension);
//   * a tcgen05.mma costs >= ~53 cycles to issue from one thread whatever its N
//     (measured, tools/microbench4/5), so the 32 MMAs of a step are issued by FOUR
//     warps into four accumulators that the gate-math warps add up (680 vs 1530 cycles);
//   * the four gates of a unit sit in four adjacent TMEM lanes; a 4x4 transpose through
//     shared memory gives each thread (unit, 4 batches, all gates); gate math in fp32,
//     c stays in registers for the whole sequence;
//   * (waits on those mbarriers use the default CTA-scope acquire, like TMA-multicast consumers:
//     a cluster-scope acquire makes ptxas emit CCTL.IVALL -- an L1 invalidate -- in the spin loop)
//   * h_t (forward) / partial dh (backward) travel across the cluster with
//     cp.async.bulk shared::cta -> shared::cluster, counted on the receivers' mbarriers
//     (no cluster barrier on the critical path), double-buffered by step parity;
//   * time-batched operands (G_t, saved gates, dy) are prefetched by TMA into rings;
//   * NCHAIN independent batch groups per cluster are interleaved so that one chain's
//     DSMEM exchange overlaps the other chain's MMA + gate math.
#include "lstm_internal.cuh"
#include "sm100.cuh"
#include "lstm_rec_tc.cuh"

namespace b2 {
using namespace sm100;

constexpr int RGS = 3;        // TMA ring stages
constexpr int NISSW = 4;      // MMA issuer warps
constexpr int ACC_STRIDE = 32; // TMEM columns between accumulators (16 used)

// Gate-math team of a chain: GW warps.  A warp can only read the TMEM lane quarter 32*(warp_id % 4), so
// with GW = 8 the two warps that share a quarter split the 16 batch columns of the accumulator:
// NBW = 16 / (GW/4) columns per warp, CPT = NBW / 4 cells per thread.  GW = 8 halves the dependent
// chain of the cell math per step (2 cells per thread instead of 4) and the per-warp share of the
// DSMEM copy issue (2 copies instead of 4).
template <int GW> struct GateTeam {
  static_assert(GW == 4 || GW == 8, "gate team = 4 or 8 warps");
  static constexpr int HF = GW / 4;
  static constexpr int NBW = 16 / HF;
  static constexpr int CPT = NBW / 4;
  static constexpr int PITCH = NBW + 4;      // transpose scratch pitch (floats)
  static constexpr int SENDS = 16 / GW;      // bulk copies issued per warp and step
};

// phase timers of the gate-math warps: compiled out unless -DB2_REC_TIMING=1
#ifndef B2_REC_TIMING
#define B2_REC_TIMING 0
#endif
#if B2_REC_TIMING
#define REC_CLK(name) const long long name = clock64()
#else
#define REC_CLK(name) do {} while (0)
#endif

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// warp-specialised register budget (whole warpgroups): the issuer / producer warps keep 24 registers,
// what they give up lets the gate-math warps run without spills
template <int N> __device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ void fence_acq_rel_cluster() {
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ======================================================================== forward
template <int NCHAIN, int GW>
struct RecSmem {
  static constexpr int kHbufOff = 0;                                   // [NCHAIN][2][32*H]  (H<=512 -> 16 KB)
  static constexpr int kHbufBytes = 32 * 512;
  static constexpr int kStageOff = kHbufOff + NCHAIN * 2 * kHbufBytes; // [NCHAIN][2][1 KB]
  static constexpr int kGOff = kStageOff + NCHAIN * 2 * 1024;          // [NCHAIN][RGS][8 KB]
  static constexpr int kScrOff = kGOff + NCHAIN * RGS * 8192;          // [NCHAIN*GW][32*PITCH*4]
  static constexpr int kBarOff = kScrOff + NCHAIN * GW * 32 * GateTeam<GW>::PITCH * 4;
  static constexpr int kBytes = kBarOff + 1024;
};

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 16) tmem_ld_32x32b_x16(taddr, r); else tmem_ld_32x32b_x8(taddr, r);
}

// warp roles: 0..3 = MMA issuers, 4 = G producer, 5.. = gate math (GW warps per chain),
// then one output-store warp per chain
template <int NCHAIN, int KS, int GW>     // KS = H/16 MMA k-steps per time step
__global__ void __launch_bounds__(160 + (GW * 32 + 32) * NCHAIN, 1)
lstm_rec_fwd_kernel(const __grid_constant__ CUtensorMap tmG, const RecFwdArgs a) {
  using L = RecSmem<NCHAIN, GW>;
  using GT = GateTeam<GW>;
  constexpr int NBW = GT::NBW, CPT = GT::CPT, PITCH = GT::PITCH;
  constexpr int NISS = KS < NISSW ? KS : NISSW;      // issuer warps actually used
  constexpr int KPER = KS / NISS;                    // k-steps per issuer
  constexpr int GTHREADS = GW * 32;                  // gate-math threads per chain
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, T = a.T, B = a.B;
  const int CS = H / RU;
  const uint32_t cta = cluster_ctarank();
  const int cluster_id = blockIdx.x / CS;
  const int dir = cluster_id & 1;
  const int gbase = (cluster_id >> 1) * NCHAIN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint64_t* bars = (uint64_t*)(smem + L::kBarOff);
  uint64_t* hfull = bars;                      // [NCHAIN][2][16]: one per (buffer, source CTA slice)
  uint64_t* accfull = bars + NCHAIN * 2 * 16;  // [NCHAIN]
  uint64_t* gfull = accfull + NCHAIN;          // [NCHAIN][RGS]
  uint64_t* gempty = gfull + NCHAIN * RGS;     // [NCHAIN][RGS]
  uint64_t* stfull = gempty + NCHAIN * RGS;    // [NCHAIN][2] staged h ready for the store warp
  uint64_t* stfree = stfull + NCHAIN * 2;      // [NCHAIN][2] store warp done with the staging buffer
  uint32_t* tmem_slot = (uint32_t*)(stfree + NCHAIN * 2);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NCHAIN * 2; ++i) { mbar_init(&stfull[i], 1); mbar_init(&stfree[i], 1); }
    for (int i = 0; i < NCHAIN * 2 * 16; ++i) mbar_init(&hfull[i], 1);
    for (int i = 0; i < NCHAIN; ++i) mbar_init(&accfull[i], NISS);
    for (int i = 0; i < NCHAIN * RGS; ++i) { mbar_init(&gfull[i], 1); mbar_init(&gempty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  // zero both h buffers of every chain (h_{-1} = 0)
  for (int i = threadIdx.x; i < NCHAIN * 2 * L::kHbufBytes / 16; i += blockDim.x)
    ((uint4*)(smem + L::kHbufOff))[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tA = tmem;                    // columns [0, H/2): the weight slice
  const uint32_t tAcc = tmem + 256;            // [NCHAIN][NISS] accumulators of 16 columns

  // ---- load this CTA's 128 x H bf16 weight slice into TMEM (first four gate-math warps: one per lane quarter)
  if (warp >= 5 && warp < 9) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint4* src = (const uint4*)(a.wpack + (((size_t)dir * CS + cta) * 128 + r) * H);
    for (int c0 = 0; c0 < H / 2; c0 += 32) {   // 32 TMEM columns = 64 bf16 = 8 x uint4
      uint32_t v[32];
      const int nvec = (H / 2 - c0) >= 32 ? 8 : (H / 2 - c0) / 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (j < nvec) u = __ldg(&src[c0 / 4 + j]);
        v[4 * j] = u.x; v[4 * j + 1] = u.y; v[4 * j + 2] = u.z; v[4 * j + 3] = u.w;
      }
      tmem_st_32x32b_x32(tA + ((uint32_t)(q * 32) << 16) + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();                              // every CTA has its barriers + zeroed buffers

  if (warp < NISSW) {
    // ------------------------------------------------------------- MMA issuers
    if (lane == 0 && warp < NISS) {
      const uint32_t idesc = make_idesc_bf16(128, RN, 0, 0);
      uint32_t hphase = 0;                       // bit (c*2+p)*4 + local slice: parity of hfull[c][p][slice]
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem + L::kHbufOff), 256, 128, 0);
      // the chains drift against each other: poll both and serve whichever has its h ready
      int tc[NCHAIN];
      int remaining = 0;
#pragma unroll
      for (int c = 0; c < NCHAIN; ++c) {
        tc[c] = (gbase + c < a.NG) ? 0 : T;
        remaining += T - tc[c];
      }
      while (remaining > 0) {
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c) {
          const int t = tc[c];
          if (t >= T) continue;
          const int p = t & 1;
          // h arrives slice by slice (one 32-unit slice per source CTA, own mbarrier each):
          // the two MMAs of a slice are issued as soon as that slice has landed, so the
          // tensor pipe works underneath the DSMEM all-gather instead of after it
          const int sl_first = (warp * KPER) >> 1;
          if (t > 0) {
            if (!mbar_try_wait(&hfull[(c * 2 + p) * 16 + sl_first], (hphase >> ((c * 2 + p) * 4)) & 1u)) continue;
            hphase ^= 1u << ((c * 2 + p) * 4);
          }
          tc_fence_after();
          // one descriptor per buffer; the k-th step only moves the start address by 512 B
          const uint64_t bd0 = bdesc0 + (uint64_t)((c * 2 + p) * (L::kHbufBytes >> 4));
          const uint32_t acc = tAcc + (c * NISS + warp) * ACC_STRIDE;
#pragma unroll
          for (int kk = 0; kk < KPER; ++kk) {
            const int k = warp * KPER + kk;
            const int ls = (k >> 1) - sl_first;          // local slice index of this issuer (0..3)
            if (t > 0 && ls > 0 && (kk == 0 || ((k - 1) >> 1) != (k >> 1))) {
              mbar_wait(&hfull[(c * 2 + p) * 16 + (k >> 1)], (hphase >> ((c * 2 + p) * 4 + ls)) & 1u);
              hphase ^= 1u << ((c * 2 + p) * 4 + ls);
              tc_fence_after();
            }
            mma_ts(acc, tA + k * 8, bd0 + (uint64_t)(k * 32), idesc, kk > 0 ? 1u : 0u);
          }
          mma_commit(&accfull[c]);
          tc[c] = t + 1;
          --remaining;
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------- G producer (TMA)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < T; ++t) {
        const int td = dir ? T - 1 - t : t;
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c) {
          if (gbase + c >= a.NG) continue;
          mbar_wait(&gempty[c * RGS + stage], phase ^ 1);
          mbar_expect_tx(&gfull[c * RGS + stage], 8192);
          tma_load_3d(smem + L::kGOff + (c * RGS + stage) * 8192, &tmG, &gfull[c * RGS + stage],
                      cta * 128, dir, td * B + (gbase + c) * RN);
        }
        if (++stage == RGS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 5 + GW * NCHAIN) {
    // ------------------------------------------------------------- output store warp
    const int c = warp - 5 - GW * NCHAIN;
    const int grp = gbase + c;
    if (grp < a.NG) {
      uint8_t* stage_base = smem + L::kStageOff + c * 2 * 1024;
      int ob[2], okc[2], olen[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ch = lane + 32 * i;             // 64 chunks of 8 units x 1 batch row
        ob[i] = grp * RN + (ch >> 2); okc[i] = ch & 3;
        olen[i] = ob[i] < B ? a.seq_len[ob[i]] : 0;
      }
      for (int t = 0; t < T; ++t) {
        const int td = dir ? T - 1 - t : t;
        const int p = t & 1;
        mbar_wait(&stfull[c * 2 + p], (t >> 1) & 1);
        const uint8_t* stg = stage_base + p * 1024;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (ob[i] >= B) continue;
          const int bl = (lane + 32 * i) >> 2;
          const uint4 raw = *(const uint4*)(stg + okc[i] * 256 + (bl >> 3) * 128 + (bl & 7) * 16);
          const bool oact = td < olen[i];
          const uint4 hv = oact ? raw : make_uint4(0, 0, 0, 0);
          const size_t o0 = ((size_t)td * B + ob[i]) * 2 * H + (size_t)dir * H + cta * RU + okc[i] * 8;
          if (a.hs_lp) *(uint4*)(a.hs_lp + o0) = hv;
          const uint32_t w[4] = {hv.x, hv.y, hv.z, hv.w};
          float f[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            f[2 * k] = __uint_as_float(w[k] << 16);
            f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
          }
          if (a.keep_prob < 1.f) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              f[k] = dropout_keep(a.seed, o0 + k, a.keep_prob) ? f[k] / a.keep_prob : 0.f;
            if (a.y_lp) {
              uint32_t pk[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 b2v = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                pk[k] = *(uint32_t*)&b2v;
              }
              *(uint4*)(a.y_lp + o0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else if (a.y_lp && a.y_lp != a.hs_lp) {
            *(uint4*)(a.y_lp + o0) = hv;
          }
          *(float4*)(a.y + o0) = make_float4(f[0], f[1], f[2], f[3]);
          *(float4*)(a.y + o0 + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&stfree[c * 2 + p]);
        if (a.progress) {
          // this direction has now covered every frame of chunk td / chunk_T (fw ascends, bw descends)
          const bool done = dir ? (td % a.chunk_T == 0) : ((td + 1) % a.chunk_T == 0 || td == T - 1);
          if (done && lane == 0) { __threadfence(); atomicAdd(a.progress + td / a.chunk_T, 1u); }
        }
      }
    }
  } else {
    // ------------------------------------------------------------- gate math
    const int c = (warp - 5) / GW;              // chain of this warp
    const int grp = gbase + c;
    if (grp < a.NG) {
      const int gw = (warp - 5) - c * GW;       // warp inside the chain's team
      const int q = warp & 3;                   // TMEM lane quarter (hardware: warp_id % 4)
      const int col0 = (gw >> 2) * NBW;         // first batch column of this warp's accumulator slice
      const int ug = lane >> 2, gq = lane & 3;
      const int ul = q * 8 + ug;                // unit inside the CTA
      const int u = cta * RU + ul;              // unit inside the layer
      const int ctid = threadIdx.x - 160 - c * GTHREADS;   // 0..GTHREADS-1 inside the chain
      float* scr = (float*)(smem + L::kScrOff) + (size_t)(warp - 5) * 32 * PITCH;
      uint8_t* stage_base = smem + L::kStageOff + c * 2 * 1024;
      int bl[CPT], bidx[CPT], len[CPT];
      float cst[CPT], hst[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        bl[j] = col0 + gq * CPT + j;            // batch column inside the chain's group of 16
        bidx[j] = grp * RN + bl[j];
        len[j] = bidx[j] < B ? a.seq_len[bidx[j]] : 0;
        cst[j] = 0.f; hst[j] = 0.f;
      }
      float pwi = 0.f, pwf = 0.f, pwo = 0.f;
      if (a.use_peephole) { pwi = a.wi[dir][u]; pwf = a.wf[dir][u]; pwo = a.wo[dir][u]; }
      size_t cell0[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) cell0[j] = ((size_t)bidx[j] * 2 + dir) * H + u;
      const size_t cell_step = (size_t)B * 2 * H;
      int stage = 0; uint32_t gph = 0;
#if B2_REC_TIMING
      const long long loop_t0 = clock64();
      unsigned long long loop_g0;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(loop_g0));
#endif
      for (int t = 0; t < T; ++t) {
        const int td = dir ? T - 1 - t : t;
        REC_CLK(e0);
        const int p = t & 1;
        // off the critical path (the MMAs of this step are still running): G_t into registers,
        // and make sure the store warp is done with the staging buffer we are about to reuse
        mbar_wait(&gfull[c * RGS + stage], gph);
        const float* Gs = (const float*)(smem + L::kGOff + (c * RGS + stage) * 8192);
        float4 G4[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) G4[j] = *(const float4*)(Gs + bl[j] * 128 + ul * 4);   // [b][u][gate]
        if (t >= 2) mbar_wait(&stfree[c * 2 + p], ((t >> 1) - 1) & 1);
        REC_CLK(e1);
        mbar_wait(&accfull[c], t & 1);
        tc_fence_after();
        REC_CLK(e2);
        float v[NBW];
        {
          uint32_t w[NISS][NBW];
#pragma unroll
          for (int s2 = 0; s2 < NISS; ++s2)
            tmem_ld_cols<NBW>(tAcc + (c * NISS + s2) * ACC_STRIDE + ((uint32_t)(q * 32) << 16) + col0, w[s2]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < NBW; ++i) {
            float acc = __uint_as_float(w[0][i]);
#pragma unroll
            for (int s2 = 1; s2 < NISS; ++s2) acc += __uint_as_float(w[s2][i]);
            v[i] = acc;
          }
        }
        // (4 gates) x (NBW batches) transpose inside each 4-lane group through shared memory:
        // lane = gate row, afterwards thread (unit, batch quad) holds all four gates of its CPT cells
#pragma unroll
        for (int j = 0; j < NBW / 4; ++j)
          *(float4*)&scr[lane * PITCH + 4 * j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        float z[4][CPT];                         // [gate][cell j]
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if constexpr (CPT == 4) {
            const float4 f = *(const float4*)&scr[(ug * 4 + g) * PITCH + gq * 4];
            z[g][0] = f.x; z[g][1] = f.y; z[g][2] = f.z; z[g][3] = f.w;
          } else {
            const float2 f = *(const float2*)&scr[(ug * 4 + g) * PITCH + gq * 2];
            z[g][0] = f.x; z[g][1] = f.y;
          }
        }
        __syncwarp();
        REC_CLK(e3);
        uint8_t* stg = stage_base + p * 1024;
        // loads first, stores last: a shared-memory store between two cells would serialise
        // them (the compiler must assume it aliases the next cell's loads)
        float4 gsv[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const bool active = td < len[j];
          const float c_prev = cst[j];
          // branch-free so that the cells of a thread interleave (ILP); inactive steps
          // (t >= seq_len) discard the result below
          float zi = z[0][j] + G4[j].x, zg = z[1][j] + G4[j].y;
          float zf = z[2][j] + G4[j].z + a.forget_bias, zo = z[3][j] + G4[j].w;
          zi = fmaf(pwi, c_prev, zi); zf = fmaf(pwf, c_prev, zf);
          // three activations share one reciprocal: 1/((1+Ei)(1+Ef)(1+Eg))
          const float Ei = __expf(fminf(-zi, 25.f)), Ef = __expf(fminf(-zf, 25.f));
          const float Eg = __expf(fminf(-2.f * zg, 25.f));
          const float ai = 1.f + Ei, af = 1.f + Ef, ag = 1.f + Eg;
          const float r = fast_rcp(ai * af * ag);
          const float gi = r * af * ag, gf = r * ai * ag, gg = (1.f - Eg) * r * ai * af;
          float c_new = fmaf(gf, c_prev, gi * gg);
          if (a.cell_clip > 0.f) c_new = fminf(fmaxf(c_new, -a.cell_clip), a.cell_clip);
          zo = fmaf(pwo, c_new, zo);
          const float Eo = __expf(fminf(-zo, 25.f)), Ec = __expf(fminf(-2.f * c_new, 25.f));
          const float ao = 1.f + Eo, ac = 1.f + Ec;
          const float r2 = fast_rcp(ao * ac);
          const float go = r2 * ac;
          const float h_out = go * (1.f - Ec) * r2 * ao;
          cst[j] = active ? c_new : c_prev;
          hst[j] = active ? h_out : hst[j];
          gsv[j] = make_float4(gi, gg, gf, go);
        }
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          // state h (carried through inactive steps) feeds the next step's GEMM
          const int off = (ul >> 3) * 256 + (bl[j] >> 3) * 128 + (bl[j] & 7) * 16 + (ul & 7) * 2;
          *(__nv_bfloat16*)(stg + off) = __float2bfloat16(hst[j]);
        }
        REC_CLK(e4);
        fence_proxy_async_smem();                 // staged h visible to the bulk-copy engine
        named_bar_sync(1 + c, GTHREADS);
        REC_CLK(e5);
        if (ctid == 0) { mbar_arrive(&gempty[c * RGS + stage]); mbar_arrive(&stfull[c * 2 + p]); }
        if (t + 1 < T && ctid < CS) mbar_expect_tx(&hfull[(c * 2 + (p ^ 1)) * 16 + ctid], 1024);
        // GT::SENDS lanes in each of the chain's GW warps issue the CS bulk copies (one per peer): a warp
        // issues one async-proxy operation per ~53 cycles, so the issue is spread over all warps of the team
        {
          const int dstcta = gw * GT::SENDS + lane;
          if (t + 1 < T && lane < GT::SENDS && dstcta < CS) {
            uint8_t* dst = smem + L::kHbufOff + (c * 2 + (p ^ 1)) * L::kHbufBytes + cta * 1024;
            bulk_s2cluster(dst, stg, 1024, &hfull[(c * 2 + (p ^ 1)) * 16 + cta], (uint32_t)dstcta);
          }
        }
        REC_CLK(e6);
        // reserve for BPTT: after the hand-off, so that these global stores overlap the DSMEM flight and
        // the next step's MMAs instead of delaying them
        if (a.gates) {
#pragma unroll
          for (int j = 0; j < CPT; ++j)
            if (bidx[j] < B) {
              const size_t cell = cell0[j] + (size_t)td * cell_step;
              *(float4*)(a.gates + cell * 4) = gsv[j];
              a.cs[cell] = cst[j];
            }
        }
#if B2_REC_TIMING
        if (a.dbg && blockIdx.x == 0 && ctid == 0 && c == 0) {
          const long long e7 = clock64();
          a.dbg[2] += e2 - e1;   // wait accumulator
          a.dbg[3] += e3 - e2;   // tmem ld + transpose
          a.dbg[4] += e1 - e0;   // G prefetch + staging-free wait (overlaps the MMAs)
          a.dbg[5] += e4 - e3;   // gate math + staging
          a.dbg[6] += e5 - e4;   // fence + named barrier
          a.dbg[7] += e6 - e5;   // arrives + sends
          a.dbg[1] += e7 - e6;   // reserve stores
        }
#endif
        if (++stage == RGS) { stage = 0; gph ^= 1; }
      }
#if B2_REC_TIMING
      if (a.dbg && ctid == 0 && cta == 0) {
        a.dbg[16 + cluster_id * 2 + c] = clock64() - loop_t0;
        unsigned long long loop_g1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(loop_g1));
        a.dbg[32 + cluster_id * 2 + c] = (long long)(loop_g1 - loop_g0);      // nanoseconds of the same loop
      }
#endif
      if (a.final_state) {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (bidx[j] < B) {
            a.final_state[((size_t)(dir * 2 + 0) * B + bidx[j]) * H + u] = cst[j];
            a.final_state[((size_t)(dir * 2 + 1) * B + bidx[j]) * H + u] = hst[j];
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                // nobody exits while peers may still write here
  if (warp == 0) tmem_dealloc(tmem, 512);
}

bool rec_tc_supported(int H) {
  if (H % RU) return false;
  const int cs = H / RU;
  return cs == 1 || cs == 2 || cs == 4 || cs == 8 || cs == 16;
}

template <int NCHAIN, int KS, int GW>
static int launch_rec_fwd(const CUtensorMap& tmG, const RecFwdArgs& a, int nclusters, int CS,
                          cudaStream_t stream) {
  using L = RecSmem<NCHAIN, GW>;
  auto kern = lstm_rec_fwd_kernel<NCHAIN, KS, GW>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
  if (CS > 8) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CS);
  cfg.blockDim = dim3(160 + (GW * 32 + 32) * NCHAIN);
  cfg.dynamicSmemBytes = L::kBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  B2_CUDA(cudaLaunchKernelEx(&cfg, kern, tmG, a));
  count_launches(1);
  return B2_OK;
}

// G: [T*B, 8H] fp32 gate pre-activations, column = dir*4H + u*4 + gate (packed order)
// gate_warps: 4 or 8 warps per chain in the gate-math team (0 = default)
int rec_tc_forward(RecFwdArgs a, const float* G, int nchain, int gate_warps, cudaStream_t stream) {
  const int H = a.H, CS = H / RU;
  a.NG = cdiv(a.B, RN);
  if (nchain < 1) nchain = (a.NG >= 2) ? 2 : 1;
  if (nchain > 2) nchain = 2;
  if (gate_warps != 4 && gate_warps != 8) gate_warps = kDefaultGateWarps;
  const int nclusters = 2 * cdiv(a.NG, nchain);
  CUtensorMap tmG;
  const uint64_t dims[3] = {(uint64_t)4 * H, 2, (uint64_t)a.T * a.B};
  const uint64_t strides[2] = {(uint64_t)4 * H * 4, (uint64_t)8 * H * 4};
  const uint32_t box[3] = {128, 1, RN};
  int rc = make_tmap_generic(&tmG, 1, G, 3, dims, strides, box, 0);
  if (rc) return rc;
#define B2_REC_DISPATCH(KS_)                                                                   \
  if (H / 16 == KS_) {                                                                        \
    if (gate_warps == 8) {                                                                    \
      if (nchain == 2) return launch_rec_fwd<2, KS_, 8>(tmG, a, nclusters, CS, stream);      \
      return launch_rec_fwd<1, KS_, 8>(tmG, a, nclusters, CS, stream);                       \
    }                                                                                         \
    if (nchain == 2) return launch_rec_fwd<2, KS_, 4>(tmG, a, nclusters, CS, stream);        \
    return launch_rec_fwd<1, KS_, 4>(tmG, a, nclusters, CS, stream);                         \
  }
  B2_REC_DISPATCH(2) B2_REC_DISPATCH(4) B2_REC_DISPATCH(8) B2_REC_DISPATCH(16) B2_REC_DISPATCH(32)
#undef B2_REC_DISPATCH
  set_error("rec_tc_forward: unsupported H=%d", H);
  return B2_ERR_UNSUPPORTED;
}

// ======================================================================== backward
// BPTT mirror.  Per step and chain:
//   A) gate-math warps: dh = dy_t + sum over the 16 peers' partial (dz_{next} . Wh^T) slices,
//      gate derivatives -> dz_t (bf16) into the local B-operand buffer and into dG (HBM);
//   B) issuer warp m: D_m[128 units x 16] = Wh[128m.., this CTA's 128 gate cols] . dz_t^T
//      (8 MMAs, A = transposed weight slice resident in TMEM);
//   C) gate-math warps: accumulators -> bf16 -> one 1 KB slice per peer (its 32 units),
//      bulk-copied into the peers' receive buffers (reduce-scatter over DSMEM).
constexpr int BGS = 2;                        // TMA ring stages (backward)
constexpr int BSTAGE = 8192 + 3 * 2048;       // gates | cs_t | cs_prev | dy

template <int NCHAIN>
struct RecBwdSmem {
  static constexpr int kRecvOff = 0;                                     // [NCHAIN][2][16 KB]
  static constexpr int kSendOff = kRecvOff + NCHAIN * 2 * 16384;         // [NCHAIN][2][16 KB]
  static constexpr int kBopOff = kSendOff + NCHAIN * 2 * 16384;          // [NCHAIN][4 KB]
  static constexpr int kRingOff = kBopOff + NCHAIN * 4096;               // [NCHAIN][BGS][BSTAGE]
  static constexpr int kBarOff = kRingOff + NCHAIN * BGS * BSTAGE;
  static constexpr int kBytes = kBarOff + 512;
};

template <int NCHAIN, int GW>
__global__ void __launch_bounds__(256 + GW * 32 * NCHAIN, 1)
lstm_rec_bwd_kernel(const __grid_constant__ CUtensorMap tmGates, const __grid_constant__ CUtensorMap tmCs,
                    const __grid_constant__ CUtensorMap tmDy, const RecBwdArgs a) {
  using L = RecBwdSmem<NCHAIN>;
  using GT = GateTeam<GW>;
  constexpr int NBW = GT::NBW, CPT = GT::CPT;
  constexpr int GTHREADS = GW * 32;
  // warp roles (warpgroup aligned, so that setmaxnreg can move registers from the issuer / producer
  // warpgroups to the gate-math teams): 0..3 MMA issuers | 4 .. 4+GW*NCHAIN-1 gate math | then the TMA
  // producer warpgroup (its first warp works, the other three only give up their registers)
  constexpr int W_GATE0 = 4, W_PROD = 4 + GW * NCHAIN;
  // launch allocation (from __launch_bounds__): 168 / 128 / 128 / 80 registers for (GW,NCHAIN) = (4,1) (4,2) (8,1) (8,2);
  // gate threads + 56 per issuer thread + 24 per producer-warpgroup thread stay inside the CTA's pool and inside
  // the 16 K registers of each SM sub-partition
  constexpr int GATE_REGS = GW == 4 ? 208 : (NCHAIN == 1 ? 128 : 96);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, T = a.T, B = a.B;
  const int CS = H / RU;
  const int MT = (H + 127) / 128;              // M tiles of 128 "unit-in" rows
  const uint32_t cta = cluster_ctarank();
  const int cluster_id = blockIdx.x / CS;
  const int dir = cluster_id & 1;
  const int gbase = (cluster_id >> 1) * NCHAIN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rall = (uint32_t)CS * 1024u;

  uint64_t* bars = (uint64_t*)(smem + L::kBarOff);
  uint64_t* rfull = bars;                      // [NCHAIN][2]  partial slices arrived
  uint64_t* bready = bars + NCHAIN * 2;        // [NCHAIN]     dz_t staged for the MMA
  uint64_t* accfull = bready + NCHAIN;         // [NCHAIN]
  uint64_t* gfull = accfull + NCHAIN;          // [NCHAIN][BGS]
  uint64_t* gempty = gfull + NCHAIN * BGS;     // [NCHAIN][BGS]
  uint32_t* tmem_slot = (uint32_t*)(gempty + NCHAIN * BGS);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NCHAIN * 2; ++i) mbar_init(&rfull[i], 1);
    for (int i = 0; i < NCHAIN; ++i) { mbar_init(&bready[i], 1); mbar_init(&accfull[i], MT); }
    for (int i = 0; i < NCHAIN * BGS; ++i) { mbar_init(&gfull[i], 1); mbar_init(&gempty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tA = tmem;                    // tile m at columns [64m, 64m+64)
  const uint32_t tAcc = tmem + 256;            // [NCHAIN][4] x 16 columns

  if (warp >= W_GATE0 && warp < W_GATE0 + 4) {   // transposed weight slice -> TMEM (one warp per lane quarter)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    for (int m = 0; m < MT; ++m) {
      const uint4* src = (const uint4*)(a.wpackT + ((((size_t)dir * CS + cta) * 4 + m) * 128 + row) * 128);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 u = __ldg(&src[half * 8 + j]);
          v[4 * j] = u.x; v[4 * j + 1] = u.y; v[4 * j + 2] = u.z; v[4 * j + 3] = u.w;
        }
        tmem_st_32x32b_x32(tA + ((uint32_t)(q * 32) << 16) + m * 64 + half * 32, v);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();
  // every CTA of this cluster is running: tell the host-side scheduler (side-stream GEMMs wait for all clusters)
  if (a.resident && cta == 0 && threadIdx.x == 0) atomicAdd(a.resident, 1u);

  if (warp < NISSW) {
    // ------------------------------------------------------------- MMA issuers (one M tile each)
    reg_dealloc<56>();
    if (lane == 0 && warp < MT) {
      const uint32_t idesc = make_idesc_bf16(128, RN, 0, 0);
      int sc[NCHAIN];
      int remaining = 0;
#pragma unroll
      for (int c = 0; c < NCHAIN; ++c) {
        sc[c] = (gbase + c < a.NG) ? 0 : T - 1;
        remaining += (T - 1) - sc[c];
      }
      while (remaining > 0) {
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c) {
          const int s = sc[c];
          if (s + 1 >= T) continue;
          if (!mbar_try_wait(&bready[c], s & 1)) continue;
          tc_fence_after();
          const uint64_t bd0 = make_smem_desc(smem_u32(smem + L::kBopOff + c * 4096), 256, 128, 0);
          const uint32_t acc = tAcc + (c * 4 + warp) * ACC_STRIDE;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma_ts(acc, tA + warp * 64 + kk * 8, bd0 + (uint64_t)(kk * 32), idesc, kk > 0 ? 1u : 0u);
          mma_commit(&accfull[c]);
          sc[c] = s + 1;
          --remaining;
        }
      }
    }
  } else if (warp >= W_PROD) {
    // ------------------------------------------------------------- TMA producer
    reg_dealloc<24>();
    if (warp == W_PROD && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      // Progress counters (chunked dX GEMM beside this kernel) are published from HERE, not from the gate-math
      // warps (one more compare + two live registers in their step loop measured +0.12 ms per layer): a gate team's
      // arrival on gempty for step s' follows its barrier of step s', which every thread passes after issuing its
      // dG stores of step s'-1.  bw (frames ascend): frame k*chunk_T - 1 is stored at s' = k*chunk_T; fw (frames
      // descend): frame k*chunk_T at s' = T - k*chunk_T.  The outermost chunk is flagged by the team after its loop.
      int flag_s = T + BGS + 1, flag_chunk = 0;
      if (a.progress) {
        if (dir) { flag_s = a.chunk_T; flag_chunk = 0; }
        else { const int k0 = (T - 1) / a.chunk_T; flag_s = T - k0 * a.chunk_T; flag_chunk = k0; }
      }
      for (int s = 0; s < T + BGS; ++s) {
        const int td = dir ? s : T - 1 - s;
        const int tp = dir ? td + 1 : td - 1;          // previous step in forward order
        const bool flag_now = (s - BGS == flag_s) && (s - BGS < T);
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c) {
          if (gbase + c >= a.NG) continue;
          const int idx = c * BGS + stage;
          mbar_wait(&gempty[idx], phase ^ 1);          // s >= BGS: the team has passed its barrier of step s - BGS
          if (flag_now) { __threadfence(); atomicAdd(a.progress + flag_chunk, 1u); }
          if (s >= T) continue;
          mbar_expect_tx(&gfull[idx], BSTAGE);
          uint8_t* st = smem + L::kRingOff + idx * BSTAGE;
          const int row0 = td * B + (gbase + c) * RN;
          const int rowp = ((tp >= 0 && tp < T) ? tp : td) * B + (gbase + c) * RN;
          tma_load_3d(st, &tmGates, &gfull[idx], cta * 128, dir, row0);
          tma_load_3d(st + 8192, &tmCs, &gfull[idx], cta * RU, dir, row0);
          tma_load_3d(st + 8192 + 2048, &tmCs, &gfull[idx], cta * RU, dir, rowp);
          tma_load_3d(st + 8192 + 4096, &tmDy, &gfull[idx], cta * RU, dir, row0);
        }
        if (flag_now) { flag_s += a.chunk_T; flag_chunk += dir ? 1 : -1; }
        if (++stage == BGS) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------- gate math
    reg_alloc<GATE_REGS>();
    const int c = (warp - W_GATE0) / GW;
    const int grp = gbase + c;
    if (grp < a.NG) {
      const int gw = (warp - W_GATE0) - c * GW;
      const int q = warp & 3;                    // TMEM lane quarter (hardware: warp_id % 4)
      const int col0 = (gw >> 2) * NBW;
      const int ug = lane >> 2, gq = lane & 3;
      const int ul = q * 8 + ug;
      const int u = cta * RU + ul;
      const int ctid = threadIdx.x - W_GATE0 * 32 - c * GTHREADS;
      int bl[CPT], bidx[CPT], len[CPT];
      float dcs[CPT];
      float gacc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // db_i, db_g, db_f, db_o, dw_i, dw_f, dw_o
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        bl[j] = col0 + gq * CPT + j;
        bidx[j] = grp * RN + bl[j];
        len[j] = bidx[j] < B ? a.seq_len[bidx[j]] : 0;
        dcs[j] = 0.f;
      }
      // gradient of the layer's final state (encoder -> decoder bridge): enters at the first
      // active BPTT step of each utterance
      // (dcs is only rewritten by active steps, so d(c_final) simply is its initial value)
      float dfh[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) dfh[j] = 0.f;
      if (a.dfinal) {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (bidx[j] < B) {
            dcs[j] = a.dfinal[((size_t)(dir * 2 + 0) * B + bidx[j]) * H + u];
            dfh[j] = a.dfinal[((size_t)(dir * 2 + 1) * B + bidx[j]) * H + u];
          }
      }
      float pwi = 0.f, pwf = 0.f, pwo = 0.f;
      if (a.use_peephole) { pwi = a.wi[dir][u]; pwf = a.wf[dir][u]; pwo = a.wo[dir][u]; }
      uint8_t* bop = smem + L::kBopOff + c * 4096;
      __nv_bfloat16* dgp[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        dgp[j] = a.dG + (size_t)bidx[j] * 8 * H + (size_t)dir * 4 * H + u * 4;
      const size_t dg_step = (size_t)B * 8 * H;
      int stage = 0; uint32_t gph = 0;
      uint32_t rph = 0;                          // bit p: parity of rfull[c][p]
      for (int s = 0; s < T; ++s) {
        const int td = dir ? s : T - 1 - s;
        const int tn = dir ? td - 1 : td + 1;    // step processed just before (BPTT order)
        const int tp = dir ? td + 1 : td - 1;    // previous step in forward order
        const int p = s & 1;
        REC_CLK(b0);
        // ---- everything that does not depend on dh first (the TMA ring runs ahead; the peers' partial sums
        //      are still in flight): saved gates / cell states / dy from the ring, tanh(c), and the linear
        //      coefficients of dh and dc in the gate derivatives
        mbar_wait(&gfull[c * BGS + stage], gph);
        REC_CLK(b1);
        const float* Rs = (const float*)(smem + L::kRingOff + (c * BGS + stage) * BSTAGE);
        const bool tp_ok = tp >= 0 && tp < T;
        float dyq[CPT], kzo[CPT], kdc[CPT], kzi[CPT], kzg[CPT], kzf[CPT], kcp[CPT], ccv[CPT], cpv[CPT];
        bool clipz[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const float4 g4 = *(const float4*)(Rs + bl[j] * 128 + ul * 4);
          const float cc = Rs[2048 + bl[j] * 32 + ul];
          const float c_prev = tp_ok ? Rs[2048 + 512 + bl[j] * 32 + ul] : 0.f;
          dyq[j] = Rs[2048 + 1024 + bl[j] * 32 + ul];
          const float gi = g4.x, gg = g4.y, gf = g4.z, go = g4.w;
          const float Ec = __expf(fminf(-2.f * cc, 25.f));
          const float tc = (1.f - Ec) * fast_rcp(1.f + Ec);
          // dzo = dh * kzo ; dc = dc_in + dh * kdc  (kdc includes the peephole path through dzo)
          kzo[j] = tc * go * (1.f - go);
          kdc[j] = fmaf(kzo[j], pwo, go * (1.f - tc * tc));
          // dzi = dc * kzi ; dzg = dc * kzg ; dzf = dc * kzf ; dc_prev = dc * kcp
          kzi[j] = gg * gi * (1.f - gi);
          kzg[j] = gi * (1.f - gg * gg);
          kzf[j] = c_prev * gf * (1.f - gf);
          kcp[j] = fmaf(kzf[j], pwf, fmaf(kzi[j], pwi, gf));
          clipz[j] = a.cell_clip > 0.f && fabsf(cc) >= a.cell_clip;
          ccv[j] = cc; cpv[j] = c_prev;
        }
        if (a.keep_prob < 1.f) {       // DropoutWrapper mask (only when the caller did not pre-mask dy); off the
          const float inv_keep = 1.f / a.keep_prob;    // dependent chain: the partial sums are still in flight
          const size_t row_base = (size_t)td * B * 2 * H + (size_t)dir * H + u;
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const size_t oidx = row_base + (size_t)bidx[j] * 2 * H;
            dyq[j] = dropout_keep(a.seed, oidx, a.keep_prob) ? dyq[j] * inv_keep : 0.f;
          }
        }
        REC_CLK(b2);
        // ---- A) dh_rec = sum of the peers' partial slices
        float dh_rec[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) dh_rec[j] = 0.f;
        if (s > 0) {
          // the partials are written by the peers' bulk copies (async proxy, complete_tx on this barrier).
          // wait_mode 0: spin with a cluster-scope acquire (ptxas puts CCTL.IVALL into the spin loop);
          // 1: spin at CTA scope, then ONE cluster-scope acquire; 2: CTA scope only (as TMA consumers do)
          // -- measured equal for 0 and 2 (6.38 ms/layer fwd+bwd), 7.76 for 1
          if (a.wait_mode == 0) {
            mbar_wait_cluster(&rfull[c * 2 + p], (rph >> p) & 1u);
          } else {
            mbar_wait(&rfull[c * 2 + p], (rph >> p) & 1u);
            if (a.wait_mode == 1) fence_acq_rel_cluster();
          }
          rph ^= 1u << p;
          const uint8_t* rb = smem + L::kRecvOff + (c * 2 + p) * 16384 + (ul * 16 + col0 + gq * CPT) * 2;
          for (int src = 0; src < CS; ++src) {
            if constexpr (CPT == 4) {
              const uint2 raw = *(const uint2*)(rb + src * 1024);
              dh_rec[0] += __uint_as_float(raw.x << 16);
              dh_rec[1] += __uint_as_float(raw.x & 0xffff0000u);
              dh_rec[2] += __uint_as_float(raw.y << 16);
              dh_rec[3] += __uint_as_float(raw.y & 0xffff0000u);
            } else {
              const uint32_t raw = *(const uint32_t*)(rb + src * 1024);
              dh_rec[0] += __uint_as_float(raw << 16);
              dh_rec[1] += __uint_as_float(raw & 0xffff0000u);
            }
          }
        }
        REC_CLK(bL);
        uint2 pkv[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const bool active = td < len[j];
          const bool nb_active = s > 0 && tn >= 0 && tn < T && tn < len[j];
          const float dh = dyq[j] + (nb_active ? dh_rec[j] : dfh[j]);
          const float dzo = dh * kzo[j];
          float dc = fmaf(dh, kdc[j], dcs[j]);
          if (clipz[j]) dc = 0.f;
          float dzi = dc * kzi[j], dzg = dc * kzg[j];
          const float dzf = dc * kzf[j];
          dcs[j] = active ? dc * kcp[j] : dcs[j];
          dzi = active ? dzi : 0.f; dzg = active ? dzg : 0.f;
          const float dzf2 = active ? dzf : 0.f, dzo2 = active ? dzo : 0.f;
          // bias and peephole gradients accumulate in registers over the whole sequence
          gacc[0] += dzi; gacc[1] += dzg; gacc[2] += dzf2; gacc[3] += dzo2;
          gacc[4] = fmaf(dzi, cpv[j], gacc[4]); gacc[5] = fmaf(dzf2, cpv[j], gacc[5]);
          gacc[6] = fmaf(dzo2, ccv[j], gacc[6]);
          __nv_bfloat162 lo = __floats2bfloat162_rn(dzi, dzg), hi = __floats2bfloat162_rn(dzf2, dzo2);
          pkv[j].x = *(uint32_t*)&lo; pkv[j].y = *(uint32_t*)&hi;
        }
        REC_CLK(bM);
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          *(uint2*)(bop + (ul >> 1) * 256 + (bl[j] >> 3) * 128 + (bl[j] & 7) * 16 + (ul & 1) * 8) = pkv[j];
        REC_CLK(b3);
        fence_proxy_async_smem();
        tc_fence_before();                         // this thread's accumulator reads of step s-1 are done
        named_bar_sync(1 + c, GTHREADS);
        if (ctid == 0) {
          mbar_arrive(&gempty[c * BGS + stage]);
          if (s + 1 < T) mbar_arrive(&bready[c]);
        }
        // dG (operand of the time-batched weight/input-gradient GEMMs) goes out after the hand-off to the
        // issuers: the stores overlap the step's MMAs instead of delaying them
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (bidx[j] < B) *(uint2*)(dgp[j] + (size_t)td * dg_step) = pkv[j];
        if (++stage == BGS) { stage = 0; gph ^= 1; }
        if (s + 1 >= T) break;
        // ---- C) partial dh of this step -> bf16 slices for the peers
        REC_CLK(b4);
        mbar_wait(&accfull[c], s & 1);
        tc_fence_after();
        REC_CLK(b5);
        uint8_t* sst = smem + L::kSendOff + (c * 2 + (p ^ 1)) * 16384;
        {
          uint32_t v[4][NBW];
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (m < MT) tmem_ld_cols<NBW>(tAcc + (c * 4 + m) * ACC_STRIDE + ((uint32_t)(q * 32) << 16) + col0, v[m]);
          tmem_ld_wait();
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int dest = m * 4 + q;
            if (m < MT && dest < CS) {
              uint32_t pk[NBW / 2];
#pragma unroll
              for (int i = 0; i < NBW / 2; ++i) {
                __nv_bfloat162 b2v = __floats2bfloat162_rn(__uint_as_float(v[m][2 * i]), __uint_as_float(v[m][2 * i + 1]));
                pk[i] = *(uint32_t*)&b2v;
              }
              uint4* d4 = (uint4*)(sst + dest * 1024 + lane * 32 + col0 * 2);
#pragma unroll
              for (int i = 0; i < NBW / 8; ++i)
                d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
          }
        }
        REC_CLK(b6);
        fence_proxy_async_smem();
        if (ctid == 0) mbar_expect_tx(&rfull[c * 2 + (p ^ 1)], rall);
        if constexpr (GW == 4) {
          // warp q staged the COMPLETE 1 KB slices of destinations q, 4+q, 8+q, 12+q (one per M tile): it sends
          // them itself after a warp-level sync -- no CTA-wide barrier between the accumulator read and the send
          __syncwarp();
          const int dstcta = lane * 4 + q;
          if (lane < 4 && lane < MT && dstcta < CS) {
            uint8_t* dst = smem + L::kRecvOff + (c * 2 + (p ^ 1)) * 16384 + cta * 1024;
            bulk_s2cluster(dst, sst + dstcta * 1024, 1024, &rfull[c * 2 + (p ^ 1)], (uint32_t)dstcta);
          }
        } else {
          named_bar_sync(1 + c, GTHREADS);         // two warps share a slice: both halves must be staged
          const int dstcta = gw * GT::SENDS + lane;
          if (lane < GT::SENDS && dstcta < CS) {
            uint8_t* dst = smem + L::kRecvOff + (c * 2 + (p ^ 1)) * 16384 + cta * 1024;
            bulk_s2cluster(dst, sst + dstcta * 1024, 1024, &rfull[c * 2 + (p ^ 1)], (uint32_t)dstcta);
          }
        }
#if B2_REC_TIMING
        if (a.dbg && blockIdx.x == 0 && ctid == 0 && c == 0) {
          const long long b7 = clock64();
          a.dbg[1] += b1 - b0;   // wait ring
          a.dbg[8] += b2 - b1;   // ring loads + dh-independent coefficients
          a.dbg[0] += bL - b2;   // wait partials + sum
          a.dbg[9] += bM - bL;   // dh-dependent math
          a.dbg[10] += b3 - bM;  // staging stores
          a.dbg[2] += b3 - b1;   // (everything between ring wait and barrier 1)
          a.dbg[3] += b4 - b3;   // fence + bar + arrive + dG stores
          a.dbg[4] += b5 - b4;   // wait MMA
          a.dbg[5] += b6 - b5;   // tmem ld + convert + stage
          a.dbg[6] += b7 - b6;   // fence + bar + sends
        }
#endif
      }
      if (a.progress) {          // the last step's stores complete the outermost chunk of this direction
        named_bar_sync(1 + c, GTHREADS);
        if (ctid == 0) { __threadfence(); atomicAdd(a.progress + (dir ? (T - 1) / a.chunk_T : 0), 1u); }
      }
      // flush the register-accumulated bias / peephole gradients (batch quads x chains x clusters)
      if (a.dbias) {
#pragma unroll
        for (int g = 0; g < 4; ++g) atomicAdd(&a.dbias[(size_t)dir * 4 * H + u * 4 + g], gacc[g]);
        if (a.use_peephole) {
          atomicAdd(&a.dwi[dir][u], gacc[4]);
          atomicAdd(&a.dwf[dir][u], gacc[5]);
          atomicAdd(&a.dwo[dir][u], gacc[6]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int NCHAIN, int GW>
static int launch_rec_bwd(const CUtensorMap& tg, const CUtensorMap& tc, const CUtensorMap& td,
                          const RecBwdArgs& a, int nclusters, int CS, cudaStream_t stream) {
  using L = RecBwdSmem<NCHAIN>;
  auto kern = lstm_rec_bwd_kernel<NCHAIN, GW>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
  if (CS > 8) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CS);
  cfg.blockDim = dim3(256 + GW * 32 * NCHAIN);
  cfg.dynamicSmemBytes = L::kBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  B2_CUDA(cudaLaunchKernelEx(&cfg, kern, tg, tc, td, a));
  count_launches(1);
  return B2_OK;
}

int rec_tc_backward(RecBwdArgs a, const float* dy, int nchain, int gate_warps, cudaStream_t stream) {
  const int H = a.H, CS = H / RU;
  a.NG = cdiv(a.B, RN);
  if (nchain < 1) nchain = (a.NG >= 2) ? 2 : 1;
  if (nchain > 2) nchain = 2;
  if (gate_warps != 4 && gate_warps != 8) gate_warps = kDefaultGateWarps;
  const int nclusters = 2 * cdiv(a.NG, nchain);
  const uint64_t TB = (uint64_t)a.T * a.B;
  CUtensorMap tg, tc, td;
  {
    const uint64_t dims[3] = {(uint64_t)4 * H, 2, TB};
    const uint64_t strides[2] = {(uint64_t)4 * H * 4, (uint64_t)8 * H * 4};
    const uint32_t box[3] = {128, 1, RN};
    int rc = make_tmap_generic(&tg, 1, a.gates, 3, dims, strides, box, 0);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)H, 2, TB};
    const uint64_t strides[2] = {(uint64_t)H * 4, (uint64_t)2 * H * 4};
    const uint32_t box[3] = {RU, 1, RN};
    int rc = make_tmap_generic(&tc, 1, a.cs, 3, dims, strides, box, 0);
    if (rc) return rc;
    rc = make_tmap_generic(&td, 1, dy, 3, dims, strides, box, 0);
    if (rc) return rc;
  }
  if (gate_warps == 8) {
    if (nchain == 2) return launch_rec_bwd<2, 8>(tg, tc, td, a, nclusters, CS, stream);
    return launch_rec_bwd<1, 8>(tg, tc, td, a, nclusters, CS, stream);
  }
  if (nchain == 2) return launch_rec_bwd<2, 4>(tg, tc, td, a, nclusters, CS, stream);
  return launch_rec_bwd<1, 4>(tg, tc, td, a, nclusters, CS, stream);
}

}  // namespace b2
