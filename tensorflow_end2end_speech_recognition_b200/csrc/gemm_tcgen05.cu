// bf16 tensor-core GEMM for sm_100a: TMA-staged operands, tcgen05.mma with the
// fp32 accumulator in TMEM, warp-specialised persistent CTAs.
//
// This is the time-batched half of the BLSTM gate GEMMs (the reference runs
// them as MatMul inside LSTMBlockCell, models/encoders/core/blstm.py:287-320):
//   forward   G  = X . Wx          (A K-major,  B K-major after weight transpose)
//   backward  dX = dG . Wx^T       (A K-major,  B K-major)
//   wgrad     dW = [X;H]^T . dG    (A MN-major, B MN-major, split-K, fp32 atomics)
// and the fully_connected output layer (models/ctc/ctc.py:216-224).
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5
// epilogue (TMEM -> registers -> smem transpose -> coalesced global stores).
// Tile 128 x BN x 64, 4..8 smem stages, two TMEM accumulators so the epilogue of
// tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "sm100.cuh"
#include <mutex>

namespace b2 {
using namespace sm100;

constexpr int GM = 128;   // tile M (= UMMA M, one TMEM lane per row)
constexpr int GK = 64;    // tile K: 64 bf16 = one 128-byte swizzle atom
constexpr int kGemmThreads = 192;

enum { EPI_STORE_F32 = 0, EPI_ATOMIC_F32 = 1, EPI_STORE_BF16 = 2 };

struct GemmArgs {
  int M, N, K;
  int ldc;
  void* C;
  const float* bias;
  float alpha;
  int m_tiles, n_tiles, k_splits, kb_per_split, kblocks;
  int epi;
  // implicit convolution over a zero-bordered NHWC buffer (vgg.cu): K block kb belongs to kernel row
  // kb / a_tap_kb; its A tile sits a_tap_rows rows further down and (kb % a_tap_kb) * 64 columns in.  0 = plain GEMM.
  int a_tap_kb, a_tap_rows;
  // EPI_STORE_F32 only: DropoutWrapper backward mask fused into the store -- C[row, col] is kept (and scaled by
  // 1/drop_keep) iff dropout_keep(drop_seed, (drop_row0 + row) * ldc + col); drop_keep >= 1: off
  float drop_keep; unsigned long long drop_seed; long long drop_row0;
};

template <int BN> struct GemmCfg {
  static constexpr int kStageA = GM * GK * 2;
  static constexpr int kStageB = BN * GK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
  static constexpr int kEpiPitch = 36;   // floats; 16-B aligned rows, conflict-free float4 access
  static constexpr int kEpiBytes = 4 * 32 * kEpiPitch * 4;
  static constexpr int kSmem = kStages * kStage + kEpiBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * Cfg::kStageA;
  float* sEpi = (float*)(smem + kStages * Cfg::kStage);
  uint64_t* bars = (uint64_t*)(smem + kStages * Cfg::kStage + Cfg::kEpiBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles = args.m_tiles * args.n_tiles;
  const int total = tiles * args.k_splits;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int ks = w / tiles, r = w % tiles;
        const int m0 = (r / args.n_tiles) * GM, n0 = (r % args.n_tiles) * BN;
        const int kb0 = ks * args.kb_per_split;
        const int kb1 = min(kb0 + args.kb_per_split, args.kblocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], Cfg::kStage);
          uint8_t* a = sA + stage * Cfg::kStageA;
          uint8_t* b = sB + stage * Cfg::kStageB;
          if (!A_MN) {
            int acol = kb * GK, arow = m0;
            if (args.a_tap_kb) {
              const int tap = kb / args.a_tap_kb;
              acol = (kb - tap * args.a_tap_kb) * GK;
              arow = m0 + tap * args.a_tap_rows;
            }
            tma_load_2d(a, &tmA, &full[stage], acol, arow);
          } else {
#pragma unroll
            for (int i = 0; i < GM / 64; ++i)
              tma_load_2d(a + i * (64 * GK * 2), &tmA, &full[stage], m0 + i * 64, kb * GK);
          }
          if (!B_MN) {
            tma_load_2d(b, &tmB, &full[stage], kb * GK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(b + i * (64 * GK * 2), &tmB, &full[stage], n0 + i * 64, kb * GK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int ks = w / tiles;
        const int kb0 = ks * args.kb_per_split;
        const int kb1 = min(kb0 + args.kb_per_split, args.kblocks);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::kStageA);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::kStageB);
#pragma unroll
          for (int k = 0; k < GK / 16; ++k) {
            // K-major: 8-row groups 1024 B apart, advance 32 B per UMMA_K inside the atom.
            // MN-major: 64-wide MN groups one box (8 KB) apart, 8-k groups 1024 B apart,
            //           advance 16 k-rows = 2048 B per UMMA_K.
            const uint64_t ad = A_MN ? make_smem_desc(a_addr + k * 2048, 64 * GK * 2, 1024, 2)
                                     : make_smem_desc(a_addr + k * 32, 16, 1024, 2);
            const uint64_t bd = B_MN ? make_smem_desc(b_addr + k * 2048, 64 * GK * 2, 1024, 2)
                                     : make_smem_desc(b_addr + k * 32, 16, 1024, 2);
            mma_ss(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          mma_commit(&empty[stage]);           // smem slot free once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&tfull[acc]);               // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3;                    // TMEM lane quarter this warp may touch
    float* st = sEpi + (warp - 2) * 32 * Cfg::kEpiPitch;
    const bool vec_ok = (args.ldc % 4 == 0) && (((uintptr_t)args.C & 15) == 0);
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int ks = w / tiles, r = w % tiles;
      const int m0 = (r / args.n_tiles) * GM, n0 = (r % args.n_tiles) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *(float4*)&st[lane * Cfg::kEpiPitch + 4 * j] =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        // lane -> (row group lane/8, 4-column group lane%8): every store instruction writes
        // 4 rows x 128 contiguous bytes
        const int cg = lane & 7;
        const int col = n0 + c * 32 + cg * 4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (args.bias && ks == 0) {
          if (col + 0 < args.N) bv.x = args.bias[col + 0];
          if (col + 1 < args.N) bv.y = args.bias[col + 1];
          if (col + 2 < args.N) bv.z = args.bias[col + 2];
          if (col + 3 < args.N) bv.w = args.bias[col + 3];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3);
          const int row = m0 + q * 32 + rr;
          if (row >= args.M || col >= args.N) continue;
          float4 o = *(const float4*)&st[rr * Cfg::kEpiPitch + cg * 4];
          o.x = args.alpha * o.x + bv.x; o.y = args.alpha * o.y + bv.y;
          o.z = args.alpha * o.z + bv.z; o.w = args.alpha * o.w + bv.w;
          const size_t idx = (size_t)row * args.ldc + col;
          const bool full4 = col + 3 < args.N;
          if (args.epi == EPI_STORE_F32) {
            if (args.drop_keep < 1.f) {
              const unsigned long long e = (unsigned long long)(args.drop_row0 + row) * (unsigned long long)args.ldc + col;
              const float sc = 1.f / args.drop_keep;
              o.x = dropout_keep(args.drop_seed, e, args.drop_keep) ? o.x * sc : 0.f;
              o.y = dropout_keep(args.drop_seed, e + 1, args.drop_keep) ? o.y * sc : 0.f;
              o.z = dropout_keep(args.drop_seed, e + 2, args.drop_keep) ? o.z * sc : 0.f;
              o.w = dropout_keep(args.drop_seed, e + 3, args.drop_keep) ? o.w * sc : 0.f;
            }
            float* p = (float*)args.C + idx;
            if (full4 && vec_ok) *(float4*)p = o;
            else {
              p[0] = o.x;
              if (col + 1 < args.N) p[1] = o.y;
              if (col + 2 < args.N) p[2] = o.z;
              if (col + 3 < args.N) p[3] = o.w;
            }
          } else if (args.epi == EPI_ATOMIC_F32) {
            float* p = (float*)args.C + idx;
            atomicAdd(p, o.x);
            if (col + 1 < args.N) atomicAdd(p + 1, o.y);
            if (col + 2 < args.N) atomicAdd(p + 2, o.z);
            if (col + 3 < args.N) atomicAdd(p + 3, o.w);
          } else {
            __nv_bfloat16* p = (__nv_bfloat16*)args.C + idx;
            if (full4 && vec_ok) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
              uint2 pk; pk.x = *(uint32_t*)&lo; pk.y = *(uint32_t*)&hi;
              *(uint2*)p = pk;
            } else {
              p[0] = __float2bfloat16(o.x);
              if (col + 1 < args.N) p[1] = __float2bfloat16(o.y);
              if (col + 2 < args.N) p[2] = __float2bfloat16(o.z);
              if (col + 3 < args.N) p[3] = __float2bfloat16(o.w);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 2-D bf16 tensor map: inner (contiguous) extent d0, outer extent d1, row pitch ld elements.
int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t ld,
                   uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available (no driver?)"); return B2_ERR_CUDA; }
  if (((uintptr_t)base & 15) || (ld * 2) % 16) {
    set_error("tensor map: base %p / pitch %llu not 16-byte aligned", base,
              (unsigned long long)ld * 2);
    return B2_ERR_INVALID;
  }
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) d0=%llu d1=%llu ld=%llu box=%ux%u", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, box0, box1);
    return B2_ERR_CUDA;
  }
  return B2_OK;
}

// generic tiled tensor map (rank <= 5), fp32 or bf16, optional 128-byte swizzle
int make_tmap_generic(CUtensorMap* tm, int dtype_is_f32, const void* base, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      int swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available (no driver?)"); return B2_ERR_CUDA; }
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = fn(tm, dtype_is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  (cuuint32_t)rank, (void*)base, d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(rank %d) failed (%d)", rank, (int)r);
    return B2_ERR_CUDA;
  }
  return B2_OK;
}

// cap on persistent GEMM CTAs (0 = all SMs): lets a GEMM on a side stream run next to a
// recurrence kernel that owns a fixed set of SMs instead of queueing behind it
static thread_local int g_cta_limit = 0;
void gemm_set_cta_limit(int n) { g_cta_limit = n; }

// one-shot epilogue dropout of the next gemm_bf16_tc (EPI_STORE_F32) launches, until reset with keep = 1
static float g_drop_keep = 1.f;
static unsigned long long g_drop_seed = 0;
static long long g_drop_row0 = 0;
void gemm_set_store_dropout(float keep, unsigned long long seed, long long row0) {
  g_drop_keep = keep; g_drop_seed = seed; g_drop_row0 = row0;
}
static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a,
                          cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    B2_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_done = true;
  }
  const int total = a.m_tiles * a.n_tiles * a.k_splits;
  int cap = num_sms();
  if (g_cta_limit > 0 && g_cta_limit < cap) cap = g_cta_limit;
  const int grid = total < cap ? total : cap;
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, kGemmThreads, Cfg::kSmem, stream>>>(tmA, tmB, a);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// C[M,N] = alpha * op(A).op(B) (+bias) on bf16 operands.
//   a_mn = 0: A is [M rows][K contiguous] (pitch lda)   1: A is [K rows][M contiguous]
//   b_mn = 0: B is [N rows][K contiguous] (pitch ldb)   1: B is [K rows][N contiguous]
//   epi: EPI_STORE_F32 | EPI_ATOMIC_F32 (C += ..., enables split-K) | EPI_STORE_BF16
int gemm_bf16_tc(int a_mn, int b_mn, int M, int N, int K, float alpha, const __nv_bfloat16* A,
                 int lda, const __nv_bfloat16* B, int ldb, void* C, int ldc, const float* bias,
                 int epi, int k_splits_hint, cudaStream_t stream) {
  B2_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_bf16_tc: bad shape %dx%dx%d", M, N, K);
  int BN = N > 128 ? 256 : N > 64 ? 128 : N > 32 ? 64 : 32;
  if (b_mn && BN < 64) BN = 64;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.C = C; g.bias = bias; g.alpha = alpha; g.epi = epi;
  g.a_tap_kb = 0; g.a_tap_rows = 0;
  g.drop_keep = (epi == EPI_STORE_F32) ? g_drop_keep : 1.f; g.drop_seed = g_drop_seed; g.drop_row0 = g_drop_row0;
  g.m_tiles = cdiv(M, GM); g.n_tiles = cdiv(N, BN);
  g.kblocks = cdiv(K, GK);
  int splits = 1;
  if (epi == EPI_ATOMIC_F32) {
    splits = k_splits_hint > 0 ? k_splits_hint : 1;
    if (k_splits_hint <= 0) {
      const int tiles = g.m_tiles * g.n_tiles;
      while (tiles * splits * 2 <= num_sms() && g.kblocks / (splits * 2) >= 8) splits *= 2;
    }
    if (splits > g.kblocks) splits = g.kblocks;
  }
  g.kb_per_split = cdiv(g.kblocks, splits);
  g.k_splits = cdiv(g.kblocks, g.kb_per_split);
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap_bf16(&tmA, A, K, M, lda, GK, GM);
  else       rc = make_tmap_bf16(&tmA, A, M, K, lda, 64, GK);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_bf16(&tmB, B, K, N, ldb, GK, BN);
  else       rc = make_tmap_bf16(&tmB, B, N, K, ldb, 64, GK);
  if (rc) return rc;
#define DISPATCH(BN_)                                                                         \
  if (BN == BN_) {                                                                            \
    if (!a_mn && !b_mn) return launch_gemm_tc<BN_, false, false>(tmA, tmB, g, stream);        \
    if (a_mn && !b_mn) return launch_gemm_tc<BN_, true, false>(tmA, tmB, g, stream);          \
    if (!a_mn && b_mn) { if (BN_ >= 64) return launch_gemm_tc<(BN_ >= 64 ? BN_ : 64), false, true>(tmA, tmB, g, stream); } \
    if (a_mn && b_mn) { if (BN_ >= 64) return launch_gemm_tc<(BN_ >= 64 ? BN_ : 64), true, true>(tmA, tmB, g, stream); }   \
  }
  DISPATCH(256) DISPATCH(128) DISPATCH(64) DISPATCH(32)
#undef DISPATCH
  set_error("gemm_bf16_tc: no kernel for BN=%d", BN);
  return B2_ERR_UNSUPPORTED;
}

// 3x3 convolution over a zero-bordered NHWC activation buffer as ONE tensor-core GEMM (vgg.cu):
//   C[M, N] = sum_{dh < taps} A_dh[M, Ktap] . B[dh*Ktap .. , N],  A_dh row r = activation row r + dh*tap_rows,
// where a row holds Ktap = kw*Cin contiguous bf16 (overlapping rows when kw = 3: row pitch Cin < Ktap -- the three
// horizontal taps of a kernel row are adjacent in memory).  A: [a_rows, pitch lda], B: [taps*Ktap rows][N contiguous].
int gemm_bf16_tc_conv(int M, int N, int Ktap, int taps, int tap_rows, const __nv_bfloat16* A, int64_t a_rows, int lda,
                      const __nv_bfloat16* B, int ldb, float* C, int ldc, cudaStream_t stream) {
  B2_CHECK_ARG(M > 0 && N >= 64 && Ktap % GK == 0 && taps >= 1, "gemm_bf16_tc_conv: bad shape M=%d N=%d Ktap=%d", M, N, Ktap);
  const int BN = N > 128 ? 256 : N > 64 ? 128 : 64;
  GemmArgs g;
  g.M = M; g.N = N; g.K = Ktap * taps; g.ldc = ldc; g.C = C; g.bias = nullptr; g.alpha = 1.f; g.epi = EPI_STORE_F32;
  g.a_tap_kb = Ktap / GK; g.a_tap_rows = tap_rows;
  g.drop_keep = 1.f; g.drop_seed = 0; g.drop_row0 = 0;
  g.m_tiles = cdiv(M, GM); g.n_tiles = cdiv(N, BN);
  g.kblocks = g.K / GK; g.kb_per_split = g.kblocks; g.k_splits = 1;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, A, (uint64_t)Ktap, (uint64_t)a_rows, (uint64_t)lda, GK, GM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, B, (uint64_t)N, (uint64_t)g.K, (uint64_t)ldb, 64, GK);
  if (rc) return rc;
  if (BN == 256) return launch_gemm_tc<256, false, true>(tmA, tmB, g, stream);
  if (BN == 128) return launch_gemm_tc<128, false, true>(tmA, tmB, g, stream);
  return launch_gemm_tc<64, false, true>(tmA, tmB, g, stream);
}

// fp32 [rows, cols] (pitch ldi) -> bf16 [rows, ldo] with zero padding of cols..ldo
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, int64_t rows, int cols, int ldi,
                                     __nv_bfloat16* __restrict__ out, int ldo) {
  const int64_t n = rows * ldo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ldo;
    const int c = (int)(i % ldo);
    out[i] = __float2bfloat16(c < cols ? in[r * ldi + c] : 0.f);
  }
}

int cast_f32_bf16(const float* in, int64_t rows, int cols, int ldi, __nv_bfloat16* out, int ldo,
                  cudaStream_t stream) {
  const int64_t n = rows * ldo;
  int blocks = (int)((n + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<blocks, 256, 0, stream>>>(in, rows, cols, ldi, out, ldo);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

int gemm_simt(int transa, int transb, int M, int N, int K, float alpha, const float* A, int lda,
              const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
              cudaStream_t stream);

}  // namespace b2

using namespace b2;

static inline int pad8(int x) { return (x + 7) / 8 * 8; }

extern "C" size_t b2_gemm_workspace_bytes(int M, int N, int K, int precision) {
  if (precision != B2_PREC_BF16) return 0;
  // bf16 copies of both operands, pitches padded to 8 elements (16 B, TMA requirement)
  const size_t a = (size_t)(M > K ? M : K) * pad8(M > K ? K : M) * 2 + (size_t)pad8(M) * pad8(K) * 2;
  const size_t b = (size_t)pad8(N > 64 ? N : 64) * pad8(K) * 2 * 2;
  return align_up(a, 256) + align_up(b, 256) + 1024;
}

static int gemm_f32_api(int transa, int transb, int M, int N, int K, float alpha, const float* A,
                        int lda, const void* A_lp, int lda_lp, const float* B, int ldb, float beta, float* C,
                        int ldc, const float* bias, int precision, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream) {
  B2_CHECK_ARG(A && B && C, "b2_gemm: null pointer");
  B2_CHECK_ARG(M > 0 && N > 0 && K > 0, "b2_gemm: bad shape %dx%dx%d", M, N, K);
  if (precision == B2_PREC_FP32)
    return gemm_simt(transa, transb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, stream);
  B2_CHECK_ARG(precision == B2_PREC_BF16, "b2_gemm: unknown precision %d", precision);
  B2_CHECK_ARG(beta == 0.f || beta == 1.f, "b2_gemm(bf16): beta must be 0 or 1");
  B2_CHECK_ARG(workspace != nullptr, "b2_gemm(bf16): workspace required");
  // operand copies: A as stored ([M,K] or [K,M]); B as stored ([K,N] -> MN-major, [N,K] -> K-major)
  const int a_rows = transa ? K : M, a_cols = transa ? M : K;
  const int b_rows = transb ? N : K, b_cols = transb ? K : N;
  const bool a_given = A_lp != nullptr && (lda_lp % 8) == 0 && (((uintptr_t)A_lp) & 15) == 0;
  const int a_ld = a_given ? lda_lp : pad8(a_cols);
  int b_ld = pad8(b_cols);
  if (!transb && b_ld < 64) b_ld = 64;        // MN-major B needs a full 64-wide box
  const size_t a_bytes = a_given ? 0 : align_up((size_t)a_rows * a_ld * 2, 256);
  const size_t b_bytes = align_up((size_t)b_rows * b_ld * 2, 256);
  if (workspace_bytes < a_bytes + b_bytes) {
    set_error("b2_gemm: workspace %zu < %zu", workspace_bytes, a_bytes + b_bytes);
    return B2_ERR_WORKSPACE;
  }
  const __nv_bfloat16* Ab = (const __nv_bfloat16*)A_lp;
  __nv_bfloat16* Bb = (__nv_bfloat16*)((char*)workspace + a_bytes);
  int rc = B2_OK;
  if (!a_given) {                              // the caller has no bf16 shadow of A: make one
    rc = cast_f32_bf16(A, a_rows, a_cols, lda, (__nv_bfloat16*)workspace, a_ld, stream);
    if (rc) return rc;
    Ab = (const __nv_bfloat16*)workspace;
  }
  rc = cast_f32_bf16(B, b_rows, b_cols, ldb, Bb, b_ld, stream);
  if (rc) return rc;
  const int epi = (beta == 1.f) ? EPI_ATOMIC_F32 : EPI_STORE_F32;
  return gemm_bf16_tc(transa ? 1 : 0, transb ? 0 : 1, M, N, K, alpha, Ab, a_ld, Bb, b_ld, C, ldc,
                      bias, epi, 0, stream);
}

extern "C" int b2_gemm(int transa, int transb, int M, int N, int K, float alpha, const float* A,
                       int lda, const float* B, int ldb, float beta, float* C, int ldc,
                       const float* bias, int precision, void* workspace, size_t workspace_bytes,
                       b2_stream_t stream_) {
  return gemm_f32_api(transa, transb, M, N, K, alpha, A, lda, nullptr, 0, B, ldb, beta, C, ldc, bias, precision,
                      workspace, workspace_bytes, (cudaStream_t)stream_);
}

extern "C" int b2_gemm_lp(int transa, int transb, int M, int N, int K, float alpha, const float* A,
                          int lda, const void* A_lp, int lda_lp, const float* B, int ldb, float beta, float* C,
                          int ldc, const float* bias, int precision, void* workspace, size_t workspace_bytes,
                          b2_stream_t stream_) {
  return gemm_f32_api(transa, transb, M, N, K, alpha, A, lda, A_lp, lda_lp, B, ldb, beta, C, ldc, bias, precision,
                      workspace, workspace_bytes, (cudaStream_t)stream_);
}

extern "C" int b2_gemm_bf16(int a_mn, int b_mn, int M, int N, int K, float alpha, const uint16_t* A,
                            int lda, const uint16_t* B, int ldb, void* C, int ldc,
                            const float* bias, int out_mode, int k_splits, b2_stream_t stream_) {
  B2_CHECK_ARG(A && B && C, "b2_gemm_bf16: null pointer");
  B2_CHECK_ARG(out_mode >= 0 && out_mode <= 2, "b2_gemm_bf16: bad out_mode %d", out_mode);
  return gemm_bf16_tc(a_mn, b_mn, M, N, K, alpha, (const __nv_bfloat16*)A, lda,
                      (const __nv_bfloat16*)B, ldb, C, ldc, bias, out_mode, k_splits,
                      (cudaStream_t)stream_);
}
