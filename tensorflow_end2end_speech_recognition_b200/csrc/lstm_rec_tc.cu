// Persistent BLSTM recurrence on tcgen05 for sm_100a (B2_PREC_BF16).
//
// The strictly sequential half of the LSTM gate GEMMs: z_t = G_t + h_{t-1} . Wh for
// T steps (reference: the LSTMBlockCell inside tf.nn.bidirectional_dynamic_rnn,
// models/encoders/core/blstm.py:287-320).  One launch runs a whole layer:
//
//   * one thread-block CLUSTER of CS = H/32 CTAs per (direction, batch group);
//     CTA `r` owns hidden units [32r, 32r+32) = 128 gate rows (unit-major, gate-minor);
//   * the CTA's slice of Wh (128 x H, bf16) is loaded ONCE into TENSOR MEMORY and
//     stays there for all T steps: the step GEMM is tcgen05.mma with A from TMEM
//     and B = h_{t-1} (16 batch columns, bf16, K-major no-swizzle) from shared memory,
//     fp32 accumulator in TMEM  (swap-AB: gates are the MMA M dimension);
//   * the four gates of one unit come out of the accumulator in four adjacent TMEM
//     lanes; a 4x4 transpose through shared memory gives each thread (unit, 4 batches,
//     all gates); gate math in fp32 with c kept in registers for the whole sequence;
//   * h_t (bf16) is all-gathered across the cluster with cp.async.bulk
//     shared::cta -> shared::cluster, completion counted on the receivers' mbarriers
//     (no cluster barrier on the critical path); double-buffered by step parity;
//   * G_t (time-batched input projection, fp32) is prefetched by TMA into a ring;
//   * NCHAIN independent batch groups per cluster are interleaved so that one chain's
//     DSMEM exchange overlaps the other chain's MMA + gate math.
#include "common.cuh"
#include "sm100.cuh"

namespace b2 {
using namespace sm100;

int make_tmap_generic(CUtensorMap* tm, int dtype_is_f32, const void* base, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      int swizzle128);
int num_sms();

constexpr int RU = 32;        // hidden units per CTA
constexpr int RN = 16;        // batch columns per chain (MMA N)
constexpr int RGS = 3;        // G ring stages
constexpr int RPITCH = 20;    // transpose scratch pitch (floats)

struct RecFwdArgs {
  int T, B, H, NG;            // NG = number of 16-wide batch groups per direction
  int D_unused;
  const int* seq_len;
  const uint16_t* wpack;      // [2][CS][128][H] bf16, row r = unit_local*4 + gate
  const float* wi[2]; const float* wf[2]; const float* wo[2];
  int use_peephole; float forget_bias, cell_clip, keep_prob; unsigned long long seed;
  float* y;                   // [T,B,2H]
  float* gates; float* cs; float* hs;   // reserve, fp32 ([T,B,2,4,H], [T,B,2,H], [T,B,2,H]) or null
  float* final_state;         // [4,B,H] or null
  long long* dbg;             // optional phase timers (clock64 sums), cluster 0 / CTA 0 only
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int NCHAIN>
struct RecSmem {
  static constexpr int kHbufOff = 0;                                   // [NCHAIN][2][32*H]  (H<=512 -> 16 KB)
  static constexpr int kHbufBytes = 32 * 512;
  static constexpr int kStageOff = kHbufOff + NCHAIN * 2 * kHbufBytes; // [NCHAIN][2][1 KB]
  static constexpr int kGOff = kStageOff + NCHAIN * 2 * 1024;          // [NCHAIN][RGS][8 KB]
  static constexpr int kScrOff = kGOff + NCHAIN * RGS * 8192;          // [NCHAIN*4][32*RPITCH*4]
  static constexpr int kBarOff = kScrOff + NCHAIN * 4 * 32 * RPITCH * 4;
  static constexpr int kBytes = kBarOff + 512;
};

// warp roles: 0 = MMA issuer, 1 = G producer, 2.. = epilogue (4 warps per chain)
template <int NCHAIN, int KS>     // KS = H/16 MMA k-steps per time step
__global__ void __launch_bounds__(64 + 128 * NCHAIN, 1)
lstm_rec_fwd_kernel(const __grid_constant__ CUtensorMap tmG, const RecFwdArgs a) {
  using L = RecSmem<NCHAIN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, T = a.T, B = a.B;
  const int CS = H / RU;
  const uint32_t cta = cluster_ctarank();
  const int cluster_id = blockIdx.x / CS;
  const int dir = cluster_id & 1;
  const int gbase = (cluster_id >> 1) * NCHAIN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t hbytes = 32u * H;             // bytes of one h buffer (16 batch x H bf16)
  const uint32_t hall = (uint32_t)CS * 1024u;  // == hbytes

  uint64_t* bars = (uint64_t*)(smem + L::kBarOff);
  uint64_t* hfull = bars;                      // [NCHAIN][2]
  uint64_t* accfull = bars + NCHAIN * 2;       // [NCHAIN]
  uint64_t* gfull = accfull + NCHAIN;          // [NCHAIN][RGS]
  uint64_t* gempty = gfull + NCHAIN * RGS;     // [NCHAIN][RGS]
  uint32_t* tmem_slot = (uint32_t*)(gempty + NCHAIN * RGS);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NCHAIN * 2; ++i) mbar_init(&hfull[i], 1);
    for (int i = 0; i < NCHAIN; ++i) mbar_init(&accfull[i], 1);
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
  const uint32_t tAcc = tmem + 256;            // 16 columns per chain

  // ---- load this CTA's 128 x H bf16 weight slice into TMEM (chain-0 epilogue warps)
  if (warp >= 2 && warp < 6) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint4* src = (const uint4*)(a.wpack + (((size_t)dir * CS + cta) * 128 + r) * H);
    for (int c0 = 0; c0 < H / 2; c0 += 32) {   // 32 TMEM columns = 64 bf16 = 8 x uint4
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 u = __ldg(&src[c0 / 4 + j]);
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

  if (warp == 0) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, RN, 0, 0);
      uint32_t hphase = 0;                       // bit (c*2+p): parity of hfull[c][p]
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem + L::kHbufOff), 256, 128, 0);
      for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c) {
          if (gbase + c >= a.NG) continue;
          const int p = t & 1;
          const long long m0 = clock64();
          if (t > 0) {
            mbar_wait_cluster(&hfull[c * 2 + p], (hphase >> (c * 2 + p)) & 1u);
            hphase ^= 1u << (c * 2 + p);
          }
          tc_fence_after();
          const long long m1 = clock64();
          // one descriptor per buffer; the k-th step only moves the start address by 512 B
          const uint64_t bd0 = bdesc0 + (uint64_t)((c * 2 + p) * (L::kHbufBytes >> 4));
#pragma unroll
          for (int k = 0; k < KS; ++k)
            mma_ts(tAcc + c * RN, tA + k * 8, bd0 + (uint64_t)(k * 32), idesc, k > 0 ? 1u : 0u);
          mma_commit(&accfull[c]);
          if (a.dbg && blockIdx.x == 0 && c == 0) {
            a.dbg[0] += m1 - m0;               // wait for h
            a.dbg[1] += clock64() - m1;        // issue 32 MMAs + commit
          }
        }
      }
    }
  } else if (warp == 1) {
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
          tma_load_4d(smem + L::kGOff + (c * RGS + stage) * 8192, &tmG, &gfull[c * RGS + stage],
                      cta * RU, 0, dir, td * B + (gbase + c) * RN);
        }
        if (++stage == RGS) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------- gate math
    const int c = (warp - 2) >> 2;              // chain of this warp
    const int grp = gbase + c;
    if (grp < a.NG) {
      const int q = warp & 3;                   // TMEM lane quarter
      const int ug = lane >> 2, gq = lane & 3;
      const int ul = q * 8 + ug;                // unit inside the CTA
      const int u = cta * RU + ul;              // unit inside the layer
      const int ctid = threadIdx.x - 64 - c * 128;   // 0..127 inside the chain
      float* scr = (float*)(smem + L::kScrOff) + (size_t)(warp - 2) * 32 * RPITCH;
      uint8_t* stage_base = smem + L::kStageOff + c * 2 * 1024;
      int bidx[4], len[4];
      float cst[4], hst[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bidx[j] = grp * RN + gq * 4 + j;
        len[j] = bidx[j] < B ? a.seq_len[bidx[j]] : 0;
        cst[j] = 0.f; hst[j] = 0.f;
      }
      float pwi = 0.f, pwf = 0.f, pwo = 0.f;
      if (a.use_peephole) { pwi = a.wi[dir][u]; pwf = a.wf[dir][u]; pwo = a.wo[dir][u]; }
      int stage = 0; uint32_t gph = 0;
      for (int t = 0; t < T; ++t) {
        const int td = dir ? T - 1 - t : t;
        const bool dbg = a.dbg && blockIdx.x == 0 && ctid == 0 && c == 0;
        const long long e0 = clock64();
        mbar_wait(&accfull[c], t & 1);
        tc_fence_after();
        const long long e1 = clock64();
        uint32_t v[16];
        tmem_ld_32x32b_x16(tAcc + c * RN + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        // 4x4 transpose inside each 4-lane group through shared memory
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *(float4*)&scr[lane * RPITCH + 4 * j] =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        float z[4][4];                           // [gate][batch j]
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 f = *(const float4*)&scr[(ug * 4 + g) * RPITCH + gq * 4];
          z[g][0] = f.x; z[g][1] = f.y; z[g][2] = f.z; z[g][3] = f.w;
        }
        __syncwarp();
        const long long e2 = clock64();
        mbar_wait(&gfull[c * RGS + stage], gph);
        const long long e3 = clock64();
        const float* Gs = (const float*)(smem + L::kGOff + (c * RGS + stage) * 8192);
        const int p = t & 1;
        uint8_t* stg = stage_base + p * 1024;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int bl = gq * 4 + j;
          const float* Gb = Gs + bl * 128 + ul;   // [b][gate][32 u]
          const bool active = td < len[j];
          const float c_prev = cst[j];
          // branch-free so that the four cells of a thread interleave (ILP); inactive steps
          // (t >= seq_len) discard the result below
          float zi = z[0][j] + Gb[0], zg = z[1][j] + Gb[32];
          float zf = z[2][j] + Gb[64] + a.forget_bias, zo = z[3][j] + Gb[96];
          zi = fmaf(pwi, c_prev, zi); zf = fmaf(pwf, c_prev, zf);
          // three activations share one reciprocal: 1/((1+Ei)(1+Ef)(1+Eg))
          const float Ei = __expf(fminf(-zi, 25.f)), Ef = __expf(fminf(-zf, 25.f));
          const float Eg = __expf(fminf(-2.f * zg, 25.f));
          const float ai = 1.f + Ei, af = 1.f + Ef, ag = 1.f + Eg;
          const float r = fast_rcp(ai * af * ag);
          float gi = r * af * ag, gf = r * ai * ag, gg = (1.f - Eg) * r * ai * af;
          float c_new = fmaf(gf, c_prev, gi * gg);
          if (a.cell_clip > 0.f) c_new = fminf(fmaxf(c_new, -a.cell_clip), a.cell_clip);
          zo = fmaf(pwo, c_new, zo);
          const float Eo = __expf(fminf(-zo, 25.f)), Ec = __expf(fminf(-2.f * c_new, 25.f));
          const float ao = 1.f + Eo, ac = 1.f + Ec;
          const float r2 = fast_rcp(ao * ac);
          float go = r2 * ac;
          float h_out = go * (1.f - Ec) * r2 * ao;
          c_new = active ? c_new : c_prev;
          h_out = active ? h_out : 0.f;
          cst[j] = c_new;
          hst[j] = active ? h_out : hst[j];
          // state h (carried through inactive steps) feeds the next step's GEMM
          const int off = (ul >> 3) * 256 + (bl >> 3) * 128 + (bl & 7) * 16 + (ul & 7) * 2;
          *(__nv_bfloat16*)(stg + off) = __float2bfloat16(hst[j]);
          if (bidx[j] < B) {
            const size_t row = (size_t)td * B + bidx[j];
            const size_t oidx = row * 2 * H + (size_t)dir * H + u;
            float yv = h_out;
            if (a.keep_prob < 1.f && active)
              yv = dropout_keep(a.seed, oidx, a.keep_prob) ? h_out / a.keep_prob : 0.f;
            a.y[oidx] = yv;
            if (a.gates) {
              float* gp = a.gates + (row * 2 + dir) * 4 * H;
              gp[u] = gi; gp[H + u] = gg; gp[2 * H + u] = gf; gp[3 * H + u] = go;
              a.cs[(row * 2 + dir) * H + u] = c_new;
              a.hs[(row * 2 + dir) * H + u] = h_out;
            }
          }
        }
        const long long e4 = clock64();
        fence_proxy_async_smem();                 // staged h visible to the bulk-copy engine
        named_bar_sync(1 + c, 128);
        const long long e5 = clock64();
        if (ctid == 0) {
          mbar_arrive(&gempty[c * RGS + stage]);
          if (t + 1 < T) mbar_expect_tx(&hfull[c * 2 + (p ^ 1)], hall);
        }
        // 4 lanes in each of the chain's 4 warps issue the CS bulk copies (one per peer):
        // spreading the issue over warps costs ~250 cycles instead of ~850 from one warp
        {
          const int dstcta = q * 4 + lane;
          if (t + 1 < T && lane < 4 && dstcta < CS) {
            uint8_t* dst = smem + L::kHbufOff + (c * 2 + (p ^ 1)) * L::kHbufBytes + cta * 1024;
            bulk_s2cluster(dst, stg, 1024, &hfull[c * 2 + (p ^ 1)], (uint32_t)dstcta);
          }
        }
        if (dbg) {
          const long long e6 = clock64();
          a.dbg[2] += e1 - e0;   // wait accumulator
          a.dbg[3] += e2 - e1;   // tmem ld + transpose
          a.dbg[4] += e3 - e2;   // wait G
          a.dbg[5] += e4 - e3;   // gate math + stores
          a.dbg[6] += e5 - e4;   // fence + named barrier
          a.dbg[7] += e6 - e5;   // issue sends
        }
        if (++stage == RGS) { stage = 0; gph ^= 1; }
      }
      if (a.final_state) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (bidx[j] < B) {
            a.final_state[((size_t)(dir * 2 + 0) * B + bidx[j]) * H + u] = cst[j];
            a.final_state[((size_t)(dir * 2 + 1) * B + bidx[j]) * H + u] = hst[j];
          }
      }
    }
  }
  (void)hbytes;
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                // nobody exits while peers may still write here
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Wh [H, 4H] fp32 (rows D.. of the TF kernel) -> bf16 [CS][128][H], row = unit_local*4 + gate
__global__ void pack_wh_kernel(const float* __restrict__ Wh, int H, uint16_t* __restrict__ out) {
  const int64_t n = (int64_t)4 * H * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % H);
    const int64_t rr = i / H;                  // cta*128 + r
    const int r = (int)(rr % 128), cta = (int)(rr / 128);
    const int ul = r >> 2, gate = r & 3;
    const __nv_bfloat16 v = __float2bfloat16(Wh[(size_t)k * 4 * H + (size_t)gate * H + cta * RU + ul]);
    out[i] = __bfloat16_as_ushort(v);
  }
}

bool rec_tc_supported(int H) {
  if (H % RU) return false;
  const int cs = H / RU;
  return cs == 1 || cs == 2 || cs == 4 || cs == 8 || cs == 16;
}

size_t rec_tc_wpack_bytes(int H) { return (size_t)2 * 4 * H * H * 2; }

int rec_tc_pack_weights(const float* kernel_fw, const float* kernel_bw, int D, int H,
                        uint16_t* wpack, cudaStream_t stream) {
  const float* k[2] = {kernel_fw, kernel_bw};
  for (int dir = 0; dir < 2; ++dir) {
    pack_wh_kernel<<<num_sms() * 4, 256, 0, stream>>>(k[dir] + (size_t)D * 4 * H, H,
                                                      wpack + (size_t)dir * 4 * H * H);
  }
  B2_LAUNCH_CHECK();
  return B2_OK;
}

template <int NCHAIN, int KS>
static int launch_rec_fwd(const CUtensorMap& tmG, const RecFwdArgs& a, int nclusters, int CS,
                          cudaStream_t stream) {
  using L = RecSmem<NCHAIN>;
  auto kern = lstm_rec_fwd_kernel<NCHAIN, KS>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
  if (CS > 8) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CS);
  cfg.blockDim = dim3(64 + 128 * NCHAIN);
  cfg.dynamicSmemBytes = L::kBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  B2_CUDA(cudaLaunchKernelEx(&cfg, kern, tmG, a));
  return B2_OK;
}

// G: [T*B, 8H] fp32 gate pre-activations (column = dir*4H + gate*H + u)
int rec_tc_forward(RecFwdArgs a, const float* G, int nchain, cudaStream_t stream) {
  const int H = a.H, CS = H / RU;
  a.NG = cdiv(a.B, RN);
  if (nchain < 1) nchain = (a.NG >= 2) ? 2 : 1;
  if (nchain > 2) nchain = 2;
  const int nclusters = 2 * cdiv(a.NG, nchain);
  CUtensorMap tmG;
  const uint64_t dims[4] = {(uint64_t)H, 4, 2, (uint64_t)a.T * a.B};
  const uint64_t strides[3] = {(uint64_t)H * 4, (uint64_t)4 * H * 4, (uint64_t)8 * H * 4};
  const uint32_t box[4] = {RU, 4, 1, RN};
  int rc = make_tmap_generic(&tmG, 1, G, 4, dims, strides, box, 0);
  if (rc) return rc;
#define B2_REC_DISPATCH(KS_)                                                          \
  if (H / 16 == KS_) {                                                               \
    if (nchain == 2) return launch_rec_fwd<2, KS_>(tmG, a, nclusters, CS, stream);  \
    return launch_rec_fwd<1, KS_>(tmG, a, nclusters, CS, stream);                   \
  }
  B2_REC_DISPATCH(2) B2_REC_DISPATCH(4) B2_REC_DISPATCH(8) B2_REC_DISPATCH(16) B2_REC_DISPATCH(32)
#undef B2_REC_DISPATCH
  set_error("rec_tc_forward: unsupported H=%d", H);
  return B2_ERR_UNSUPPORTED;
}

}  // namespace b2
