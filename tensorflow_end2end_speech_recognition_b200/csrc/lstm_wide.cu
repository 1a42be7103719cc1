// Grid-resident BLSTM recurrence for wide layers (512 < H <= 1024, BASELINE config 4: 6 x 1024) on sm_100a.
//
// The cluster/TMEM recurrence of lstm_rec_tc.cu keeps one direction's recurrent weights inside a 16-CTA cluster; at
// H = 1024 the bf16 matrix is 8 MB and does not fit, and the fallback (one split-K skinny GEMM + one gate kernel PER
// FRAME, lstm.cu) costs ~45 us per frame, L2-bound on 32 MB of fp32 weights and launch-bound.  Here the whole layer
// forward is ONE cooperative launch (same op as before: LSTMBlockCell under tf.nn.bidirectional_dynamic_rnn,
// models/encoders/core/blstm.py:287-320):
//
//   * grid = 2 directions x H/16 CTAs, all co-resident (cooperative launch; <= 148 SMs);
//   * CTA j of a direction owns 16 hidden units = 64 gate rows; its bf16 slice of Wh (64 x H = 128 KB at H = 1024)
//     lives in REGISTERS for the whole sequence as mma.sync A fragments (8 warps = 4 gates x 2 K halves,
//     H/8 = 128 registers per thread) -- the per-step GEMM [64 x H] . [H x B] streams only h_{t-1} (B <= 32 rows of
//     bf16, 64 KB) from shared memory;
//   * h_t travels through L2: every CTA writes its 16 columns of the [B, H] bf16 exchange buffer (double-buffered by
//     step parity), then one grid barrier per direction and step (release: __threadfence + atomicAdd; acquire:
//     ld.acquire spin) -- about 2 us, the cost this design accepts in exchange for having all SMs hold weights;
//   * the two K-half partial sums meet in shared memory, then 256 threads do the gate math for 16 units x 32 batch rows
//     (2 cells per thread, fp32, cell state in registers), emit y / the fp32 reserve the BPTT kernels read.
//
// tcgen05 is not used here on purpose: the step GEMM is M = 64 rows per SM with the weights stationary; with A in TMEM
// the 64 x 1024 slice would take all 512 columns and leave no accumulator, and an SS-mode MMA would re-read 128 KB of
// weights from shared memory every step.  mma.sync with register-resident A reads only h.
#include "lstm_internal.cuh"
#include "sm100.cuh"
#include <cooperative_groups.h>

namespace b2 {
using namespace sm100;

constexpr int WU = 16;            // units per CTA
constexpr int WB = 32;            // batch rows per launch (mma N = 4 tiles of 8)
constexpr int WTHREADS = 256;
constexpr int WPP = 34;           // pitch of the partial-sum tiles (floats)

struct WideFwdArgs {
  int T, B, D_in, H;
  int use_peephole; float forget_bias, cell_clip, keep_prob; unsigned long long seed;
  const float* kernel[2]; const float* wi[2]; const float* wf[2]; const float* wo[2];
  const int* seq_len;
  const float* G;                 // [T*B, 8H] fp32, column = dir*4H + gate*H + u (TF order), bias included
  float* y;                       // [T,B,2H]
  float* gates; float* cs;        // fp32 reserve ([T,B,2,H,4], [T,B,2,H]) or null
  __nv_bfloat16* hs_lp;           // [T*B, 2H] bf16 h before dropout (A operand of the dWh GEMM) or null
  __nv_bfloat16* y_lp;            // [T*B, 2H] bf16 emitted output (operand of the next layer's GEMMs) or null; may alias hs_lp
  float* final_state;             // [4,B,H] (c_fw, h_fw, c_bw, h_bw) or null
  __nv_bfloat16* hx;              // exchange buffer [2 dir][2 parity][WB][H]
  unsigned* bar;                  // [2] grid-barrier counters (zeroed before the launch)
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&v;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// KSH = k-steps (of 16) per K half = H / 32
template <int KSH>
__global__ void __launch_bounds__(WTHREADS, 1)
lstm_wide_fwd_kernel(const WideFwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = a.H, B = a.B, T = a.T;
  const int HP = H + 8;                                   // bf16 row pitch of the h tile: conflict-free fragment loads
  __nv_bfloat16* hbuf = (__nv_bfloat16*)smem_raw;          // [WB][HP]
  float* part = (float*)(smem_raw + (size_t)WB * HP * 2);  // [2 K halves][64 rows][WPP]
  uint64_t* hfull = (uint64_t*)(smem_raw + (size_t)WB * HP * 2 + (size_t)2 * 64 * WPP * 4);   // [2]: one per K half
  if (threadIdx.x == 0) { mbar_init(hfull, 1); mbar_init(hfull + 1, 1); fence_mbar_init(); }
  __syncthreads();
  const int NS = H / WU;
  const int dir = blockIdx.x / NS, slice = blockIdx.x % NS;
  const int u0 = slice * WU;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = warp & 3, kh = warp >> 2;                 // m-tile = gate (16 unit rows), K half
  const int fr = lane >> 2, fc = lane & 3;                 // fragment row / column-pair index

  // ---- this warp's weight fragments: A[row = unit][k] = Wh[k][gate*H + u0 + unit], gate = mt, k in this K half
  uint32_t afrag[KSH][4];
  {
    const float* col0 = a.kernel[dir] + (size_t)a.D_in * 4 * H + (size_t)mt * H + u0;       // + k*4H + unit
#pragma unroll
    for (int ks = 0; ks < KSH; ++ks) {
      const int k0 = kh * (KSH * 16) + ks * 16 + 2 * fc;
      afrag[ks][0] = pack_bf16(col0[(size_t)k0 * 4 * H + fr], col0[(size_t)(k0 + 1) * 4 * H + fr]);
      afrag[ks][1] = pack_bf16(col0[(size_t)k0 * 4 * H + fr + 8], col0[(size_t)(k0 + 1) * 4 * H + fr + 8]);
      afrag[ks][2] = pack_bf16(col0[(size_t)(k0 + 8) * 4 * H + fr], col0[(size_t)(k0 + 9) * 4 * H + fr]);
      afrag[ks][3] = pack_bf16(col0[(size_t)(k0 + 8) * 4 * H + fr + 8], col0[(size_t)(k0 + 9) * 4 * H + fr + 8]);
    }
  }
  // ---- gate-math ownership: unit ul = tid & 15, batch rows bq and bq + 16
  const int ul = tid & 15, bq = tid >> 4;
  const int u = u0 + ul;
  int len[2]; float cst[2], hst[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int b = bq + 16 * j;
    len[j] = b < B ? a.seq_len[b] : 0;
    cst[j] = 0.f; hst[j] = 0.f;
  }
  float pwi = 0.f, pwf = 0.f, pwo = 0.f;
  if (a.use_peephole) { pwi = a.wi[dir][u]; pwf = a.wf[dir][u]; pwo = a.wo[dir][u]; }
  __nv_bfloat16* hx_dir = a.hx + (size_t)dir * 2 * WB * H;
  unsigned* bar = a.bar + dir * 64;           // one flag per CTA of this direction: steps it has published
#if B2_REC_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define WCLK(x) const long long x = clock64()
#else
#define WCLK(x)
#endif

  for (int s = 0; s < T; ++s) {
    const int td = dir ? T - 1 - s : s;
    WCLK(w0);
    // G_t for this thread's two cells (independent of h: in flight while the barrier is awaited)
    float z[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = bq + 16 * j;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        z[j][g] = b < B ? __ldg(a.G + ((size_t)td * B + b) * 8 * H + (size_t)dir * 4 * H + (size_t)g * H + u) : 0.f;
    }

    if (s > 0) {
      // ---- every CTA of this direction has published h_{s-1}
      // h_{s-1} [B, H] bf16 from L2 into the padded tile: 2 x 32 TMA bulk row copies (one mbarrier per K half, so the
      // warps of the first half start while the second is still landing).  A warp issues one bulk operation per ~53
      // cycles whatever the lane (64 copies from one warp measured 4 300 cycles until the tile had landed): every
      // warp issues 8 of them.  The tile was last read before the block barrier that closed the previous step.
      if (warp == 0) {
        // every CTA of this direction has published step s-1: lane i polls the flags of CTAs i and i + 32 (own words
        // of one 256-byte block; no serialised atomics on a shared counter)
        for (int j = lane; j < NS; j += 32)
          while (ld_acquire_u32(bar + j) < (unsigned)s) {}
        __syncwarp();
        if (lane == 0) { mbar_expect_tx(hfull, (uint32_t)WB * H); mbar_expect_tx(hfull + 1, (uint32_t)WB * H); }
#if B2_REC_TIMING
        tacc[0] += clock64() - w0;                              // flags acquired
#endif
      }
      __syncthreads();
      if (lane < 8) {
        asm volatile("fence.proxy.async;" ::: "memory");      // the peers' generic-proxy stores, acquired by warp 0
        const int r = warp * 4 + (lane >> 1), half = lane & 1;
        bulk_g2s(hbuf + (size_t)r * HP + half * (H / 2),
                 hx_dir + (size_t)((s - 1) & 1) * WB * H + (size_t)r * H + half * (H / 2), (uint32_t)H, hfull + half);
      }
      mbar_wait(hfull + kh, (uint32_t)(s - 1) & 1u);
      WCLK(w1);
      // ---- z_rec[64 x 32] = Wslice . h^T : this warp = the 16 rows of gate mt x K half kh x all 32 batch columns
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KSH; ++ks) {
        const int k0 = kh * (KSH * 16) + ks * 16 + 2 * fc;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const __nv_bfloat16* hp = hbuf + (size_t)(nt * 8 + fr) * HP + k0;
          mma_bf16_16816(acc[nt], afrag[ks], *(const uint32_t*)hp, *(const uint32_t*)(hp + 8));
        }
      }
      WCLK(w2);
      // partial sums of this K half: part[kh][gate*16 + unit][batch]
      {
        float* pr = part + ((size_t)kh * 64 + mt * 16) * WPP;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          *(float2*)(pr + (size_t)fr * WPP + nt * 8 + 2 * fc) = make_float2(acc[nt][0], acc[nt][1]);
          *(float2*)(pr + (size_t)(fr + 8) * WPP + nt * 8 + 2 * fc) = make_float2(acc[nt][2], acc[nt][3]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int g = 0; g < 4; ++g)
          z[j][g] += part[((size_t)g * 16 + ul) * WPP + bq + 16 * j] +
                     part[((size_t)64 + g * 16 + ul) * WPP + bq + 16 * j];
#if B2_REC_TIMING
      const long long w3 = clock64();
      tacc[1] += w1 - w0; tacc[2] += w2 - w1; tacc[3] += w3 - w2;
#endif
    }
    WCLK(w4);
    // ---- gate math (models/recurrent/layers/lstm.py:142-183: i, g(ci), f, o; forget bias; peepholes; clip); the three
    //      input-side activations share one reciprocal, as in lstm_rec_tc.cu
    __nv_bfloat16* hx_out = hx_dir + (size_t)(s & 1) * WB * H;
    float4 gsv[2]; float hov[2]; bool act[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = bq + 16 * j;
      const bool active = td < len[j];
      const float c_prev = cst[j];
      float zi = z[j][0], zg = z[j][1], zf = z[j][2] + a.forget_bias, zo = z[j][3];
      zi = fmaf(pwi, c_prev, zi); zf = fmaf(pwf, c_prev, zf);
      const float Ei = __expf(fminf(-zi, 25.f)), Ef = __expf(fminf(-zf, 25.f)), Eg = __expf(fminf(-2.f * zg, 25.f));
      const float ai = 1.f + Ei, af = 1.f + Ef, ag = 1.f + Eg;
      const float r = rcp_approx(ai * af * ag);
      float gi = r * af * ag, gf = r * ai * ag, gg = (1.f - Eg) * r * ai * af;
      float c_new = fmaf(gf, c_prev, gi * gg);
      if (a.cell_clip > 0.f) c_new = fminf(fmaxf(c_new, -a.cell_clip), a.cell_clip);
      zo = fmaf(pwo, c_new, zo);
      const float Eo = __expf(fminf(-zo, 25.f)), Ec = __expf(fminf(-2.f * c_new, 25.f));
      const float ao = 1.f + Eo, ac = 1.f + Ec;
      const float r2 = rcp_approx(ao * ac);
      float go = r2 * ac;
      float h_out = go * (1.f - Ec) * r2 * ao;
      if (active) { cst[j] = c_new; hst[j] = h_out; }
      else { gi = 0.f; gg = 0.f; gf = 0.f; go = 0.f; h_out = 0.f; }
      gsv[j] = make_float4(gi, gg, gf, go); hov[j] = h_out; act[j] = active;
      // the carried state h feeds the next step's product (inactive rows keep their last state)
      hx_out[(size_t)b * H + u] = __float2bfloat16(hst[j]);
    }
    // ---- publish: block barrier (every thread's h stores precede it), then ONE release store per CTA.  The reserve
    //      stores come AFTER it: the release has to wait only for the 1 KB of h, not for the 40 KB of gates / y
    WCLK(w5);
    __syncthreads();
    WCLK(w6);
    if (tid == 0) st_release_u32(bar + slice, (unsigned)(s + 1));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = bq + 16 * j;
      if (b < B) {
        const size_t row = (size_t)td * B + b;
        const size_t oidx = row * 2 * H + (size_t)dir * H + u;
        float yv = hov[j];
        if (a.keep_prob < 1.f && act[j]) yv = dropout_keep(a.seed, oidx, a.keep_prob) ? hov[j] / a.keep_prob : 0.f;
        a.y[oidx] = yv;
        if (a.hs_lp) a.hs_lp[oidx] = __float2bfloat16(hov[j]);
        if (a.y_lp && a.y_lp != a.hs_lp) a.y_lp[oidx] = __float2bfloat16(yv);
        if (a.gates) {
          const size_t cell = (row * 2 + dir) * H + u;
          *(float4*)(a.gates + cell * 4) = gsv[j];
          a.cs[cell] = cst[j];
        }
      }
    }
#if B2_REC_TIMING
    tacc[4] += w5 - w4; tacc[5] += w6 - w5; tacc[6] += clock64() - w6; tacc[7] += clock64() - w0;
#endif
  }
#if B2_REC_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 128)) {
    long long* dbg = (long long*)(a.bar + 128) + (tid ? 8 : 0);
    for (int i = 0; i < 8; ++i) dbg[i] = tacc[i];
  }
#endif
  if (a.final_state) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = bq + 16 * j;
      if (b < B) {
        a.final_state[((size_t)(dir * 2 + 0) * B + b) * H + u] = cst[j];
        a.final_state[((size_t)(dir * 2 + 1) * B + b) * H + u] = hst[j];
      }
    }
  }
}

// ======================================================================== BPTT
// dh_t[b, u] = dy_t + sum_k dz_{t'}[b, k] . Wh[u, k]  (t' = the frame processed one BPTT step earlier, k over the 4H
// gate columns).  CTA j of a direction owns 16 units: its 16 x 4H bf16 rows of Wh are register-resident (8 warps = 8
// K eighths, H/8 registers per thread); the gate gradients dz_{t'} of ALL units ([B, 4H] bf16, 256 KB at H = 1024) are
// all-gathered through L2 behind the same grid barrier and streamed through shared memory one gate (H columns) at a
// time; the eight partial sums meet in shared memory; the gate-derivative math runs on 2 cells per thread and emits
// dz_t to the exchange buffer (bf16) and to dG (fp32, TF column order) for the time-batched GEMMs of lstm.cu.
struct WideBwdArgs {
  int T, B, D_in, H;
  int use_peephole; float cell_clip, keep_prob; unsigned long long seed;
  const float* kernel[2]; const float* wi[2]; const float* wf[2]; const float* wo[2];
  const int* seq_len;
  const float* dy;                // [T,B,2H]
  const float* gates; const float* cs;
  const float* dfinal;            // [4,B,H] or null
  float* dG;                      // [T*B, 8H] fp32, column = dir*4H + gate*H + u
  __nv_bfloat16* dG_lp;           // the same in bf16 (operand of the time-batched GEMMs) or null
  __nv_bfloat16* dzx;             // exchange buffer [2 dir][2 parity][WB][4H]
  unsigned* bar;                  // [2]
};

// KS8 = k-steps (of 16) per (gate, warp) = H / 128
template <int KS8>
__global__ void __launch_bounds__(WTHREADS, 1)
lstm_wide_bwd_kernel(const WideBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = a.H, B = a.B, T = a.T;
  const int HP = H + 8;
  // three tiles [WB][HP] bf16, each one gate's H columns of dz_{t'}; gate n of the kernel's running count lands in
  // tile n % 3 (TMA bulk row copies, completion on full[n % 3]); tile reuse is released through empty[]
  const size_t tile_bytes = (size_t)WB * HP * 2;
  float* part = (float*)(smem_raw + 3 * tile_bytes);       // [8 warps][16 units][WPP]
  uint64_t* full = (uint64_t*)(smem_raw + 3 * tile_bytes + (size_t)8 * 16 * WPP * 4);
  uint64_t* empty = full + 3;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 8); }
    fence_mbar_init();
  }
  __syncthreads();
  const int NS = H / WU;
  const int dir = blockIdx.x / NS, slice = blockIdx.x % NS;
  const int u0 = slice * WU;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fc = lane & 3;

  // A[row = unit][k] = Wh[u0 + unit][k]: rows of the TF kernel are contiguous in k
  uint32_t afrag[4][KS8][4];
  {
    const float* W0 = a.kernel[dir] + (size_t)(a.D_in + u0) * 4 * H;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int ks = 0; ks < KS8; ++ks) {
        const int k0 = g * H + warp * (KS8 * 16) + ks * 16 + 2 * fc;
        const float* r0 = W0 + (size_t)fr * 4 * H + k0;
        const float* r1 = W0 + (size_t)(fr + 8) * 4 * H + k0;
        afrag[g][ks][0] = pack_bf16(r0[0], r0[1]);
        afrag[g][ks][1] = pack_bf16(r1[0], r1[1]);
        afrag[g][ks][2] = pack_bf16(r0[8], r0[9]);
        afrag[g][ks][3] = pack_bf16(r1[8], r1[9]);
      }
  }
  const int ul = tid & 15, bq = tid >> 4;
  const int u = u0 + ul;
  int len[2]; float dcs[2], dfh[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int b = bq + 16 * j;
    len[j] = b < B ? a.seq_len[b] : 0;
    dcs[j] = 0.f; dfh[j] = 0.f;
    if (a.dfinal && b < B) {
      dcs[j] = a.dfinal[((size_t)(dir * 2 + 0) * B + b) * H + u];
      dfh[j] = a.dfinal[((size_t)(dir * 2 + 1) * B + b) * H + u];
    }
  }
  float pwi = 0.f, pwf = 0.f, pwo = 0.f;
  if (a.use_peephole) { pwi = a.wi[dir][u]; pwf = a.wf[dir][u]; pwo = a.wo[dir][u]; }
  __nv_bfloat16* zx_dir = a.dzx + (size_t)dir * 2 * WB * 4 * H;
  unsigned* bar = a.bar + dir * 64;           // one flag per CTA of this direction: steps it has published

  for (int s = 0; s < T; ++s) {
    const int td = dir ? s : T - 1 - s;
    const int tn = dir ? td - 1 : td + 1;      // frame processed one BPTT step earlier
    const int tp = dir ? td + 1 : td - 1;      // previous frame in forward order
    // ---- reserve + dy of this thread's two cells (independent of dh: in flight while the barrier is awaited)
    float4 g4[2]; float cc[2], cp[2], dyv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = bq + 16 * j;
      g4[j] = make_float4(0.f, 0.f, 0.f, 0.f); cc[j] = 0.f; cp[j] = 0.f; dyv[j] = 0.f;
      if (b < B) {
        const size_t row = (size_t)td * B + b;
        const size_t cell = (row * 2 + dir) * H + u;
        g4[j] = __ldg((const float4*)(a.gates + cell * 4));
        cc[j] = __ldg(a.cs + cell);
        if (tp >= 0 && tp < T && tp < len[j]) cp[j] = __ldg(a.cs + (((size_t)tp * B + b) * 2 + dir) * H + u);
        const size_t oidx = row * 2 * H + (size_t)dir * H + u;
        dyv[j] = __ldg(a.dy + oidx);
        if (a.keep_prob < 1.f) dyv[j] = dropout_keep(a.seed, oidx, a.keep_prob) ? dyv[j] / a.keep_prob : 0.f;
      }
    }
    float dh_rec[2] = {0.f, 0.f};
    if (s > 0) {
      const __nv_bfloat16* zsrc = zx_dir + (size_t)((s - 1) & 1) * WB * 4 * H;
      const unsigned n0 = (unsigned)(s - 1) * 4u;            // running gate-tile count at this step's first tile
      if (warp == 0) {
        // every CTA of this direction has published step s-1: lane i polls the flags of CTAs i and i + 32
        for (int j = lane; j < NS; j += 32)
          while (ld_acquire_u32(bar + j) < (unsigned)s) {}
      }
      __syncthreads();
      // every warp issues 4 of a tile's 32 row copies (a warp issues one bulk operation per ~53 cycles whatever the
      // lane: 32 from one warp take 1 700 cycles per tile)
      if (lane < 4) asm volatile("fence.proxy.async;" ::: "memory");   // the peers' generic-proxy stores, acquired above
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const unsigned n = n0 + g, tl = n % 3u;
        if (n >= 3u) mbar_wait(empty + tl, ((n / 3u) - 1u) & 1u);      // all 8 warps are done with tile use n-3
        if (tid == 0) mbar_expect_tx(full + tl, (uint32_t)WB * H * 2);
        if (lane < 4) {
          const int r = warp * 4 + lane;
          bulk_g2s(smem_raw + tl * tile_bytes + (size_t)r * HP * 2, zsrc + (size_t)r * 4 * H + (size_t)g * H,
                   (uint32_t)H * 2, full + tl);
        }
      }
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const unsigned n = n0 + g, tl = n % 3u;
        if (g == 1) {
          // the fourth gate of this step reuses the tile of the first: wait until every warp has released it
          const unsigned n3 = n0 + 3u, t3 = n3 % 3u;
          mbar_wait(empty + t3, ((n3 / 3u) - 1u) & 1u);
          if (tid == 0) mbar_expect_tx(full + t3, (uint32_t)WB * H * 2);
          if (lane < 4) {
            const int r = warp * 4 + lane;
            bulk_g2s(smem_raw + t3 * tile_bytes + (size_t)r * HP * 2, zsrc + (size_t)r * 4 * H + (size_t)3 * H,
                     (uint32_t)H * 2, full + t3);
          }
        }
        mbar_wait(full + tl, (n / 3u) & 1u);
        const __nv_bfloat16* zbuf = (const __nv_bfloat16*)(smem_raw + tl * tile_bytes);
#pragma unroll
        for (int ks = 0; ks < KS8; ++ks) {
          const int k0 = warp * (KS8 * 16) + ks * 16 + 2 * fc;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const __nv_bfloat16* zp = zbuf + (size_t)(nt * 8 + fr) * HP + k0;
            mma_bf16_16816(acc[nt], afrag[g][ks], *(const uint32_t*)zp, *(const uint32_t*)(zp + 8));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + tl);
      }
      {
        float* pr = part + (size_t)warp * 16 * WPP;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          *(float2*)(pr + (size_t)fr * WPP + nt * 8 + 2 * fc) = make_float2(acc[nt][0], acc[nt][1]);
          *(float2*)(pr + (size_t)(fr + 8) * WPP + nt * 8 + 2 * fc) = make_float2(acc[nt][2], acc[nt][3]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += part[((size_t)w * 16 + ul) * WPP + bq + 16 * j];
        dh_rec[j] = sum;
      }
    }
    // ---- gate derivatives (the fp32 step kernel's math, lstm.cu::lstm_bwd_step_kernel)
    __nv_bfloat16* zx_out = zx_dir + (size_t)(s & 1) * WB * 4 * H;
    float4 dzv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = bq + 16 * j;
      const bool active = td < len[j];
      float dzi = 0.f, dzg = 0.f, dzf = 0.f, dzo = 0.f;
      if (active) {
        const bool nb_active = s > 0 && tn >= 0 && tn < T && tn < len[j];
        const float dh = dyv[j] + (nb_active ? dh_rec[j] : dfh[j]);
        const float gi = g4[j].x, gg = g4[j].y, gf = g4[j].z, go = g4[j].w;
        const float Ec = __expf(fminf(-2.f * cc[j], 25.f));
        const float tc = (1.f - Ec) * rcp_approx(1.f + Ec);
        dzo = dh * tc * go * (1.f - go);
        float dc = dcs[j] + dh * go * (1.f - tc * tc);
        dc = fmaf(dzo, pwo, dc);
        if (a.cell_clip > 0.f && fabsf(cc[j]) >= a.cell_clip) dc = 0.f;
        dzi = dc * gg * gi * (1.f - gi);
        dzg = dc * gi * (1.f - gg * gg);
        dzf = dc * cp[j] * gf * (1.f - gf);
        dcs[j] = fmaf(dzf, pwf, fmaf(dzi, pwi, dc * gf));
      }
      __nv_bfloat16* zr = zx_out + (size_t)b * 4 * H + u;
      zr[0] = __float2bfloat16(dzi); zr[H] = __float2bfloat16(dzg);
      zr[2 * H] = __float2bfloat16(dzf); zr[3 * H] = __float2bfloat16(dzo);
      dzv[j] = make_float4(dzi, dzg, dzf, dzo);
    }
    // publish first (the release waits only for the exchanged bf16 values), the fp32 dG stores afterwards
    __syncthreads();
    if (tid == 0) st_release_u32(bar + slice, (unsigned)(s + 1));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = bq + 16 * j;
      if (b < B) {
        const size_t o = ((size_t)td * B + b) * 8 * H + (size_t)dir * 4 * H + u;
        float* dz = a.dG + o;
        dz[0] = dzv[j].x; dz[H] = dzv[j].y; dz[2 * H] = dzv[j].z; dz[3 * H] = dzv[j].w;
        if (a.dG_lp) {
          __nv_bfloat16* dl = a.dG_lp + o;
          dl[0] = __float2bfloat16(dzv[j].x); dl[H] = __float2bfloat16(dzv[j].y);
          dl[2 * H] = __float2bfloat16(dzv[j].z); dl[3 * H] = __float2bfloat16(dzv[j].w);
        }
      }
    }
  }
}

bool wide_rec_supported(const b2_lstm_desc* d) {
  if (d->precision != B2_PREC_BF16 || d->num_proj > 0) return false;
  if (!env_int("B2_WIDE_REC", 1)) return false;
  const int H = d->H;
  if (H <= 512 || H > 1024 || (H % 128) != 0) return false;          // KSH = H/32 in {20,24,28,32}: 640 768 896 1024
  if (d->B > WB) return false;
  if (2 * (H / WU) > num_sms()) return false;
  return b2_device_is_sm100() == 1;
}

size_t wide_rec_workspace_bytes(const b2_lstm_desc* d) {
  // exchange buffer (forward: h [2][2][WB][H]; BPTT: dz [2][2][WB][4H]) + barrier counters
  return align_up((size_t)2 * 2 * WB * 4 * d->H * 2, 256) + 1024;
}

template <int KSH>
static int launch_wide_fwd(WideFwdArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)WB * (a.H + 8) * 2 + (size_t)2 * 64 * WPP * 4 + 64;
  auto kern = lstm_wide_fwd_kernel<KSH>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {(void*)&a};
  B2_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(2 * (a.H / WU)), dim3(WTHREADS), params, smem, stream));
  count_launches(1);
  return B2_OK;
}

// workspace: wide_rec_workspace_bytes(d) bytes (exchange buffer | counters)
int wide_rec_forward(const b2_lstm_desc* d, const b2_lstm_params* fw, const b2_lstm_params* bw, const int32_t* seq_len,
                     const float* G, float* y, float* gates, float* cs, __nv_bfloat16* hs_lp, __nv_bfloat16* y_lp,
                     float* final_state, void* workspace,
                     cudaStream_t stream) {
  WideFwdArgs a;
  a.T = d->T; a.B = d->B; a.D_in = d->D_in; a.H = d->H;
  a.use_peephole = d->use_peephole; a.forget_bias = d->forget_bias; a.cell_clip = d->cell_clip;
  a.keep_prob = d->keep_prob; a.seed = d->dropout_seed;
  const b2_lstm_params* P[2] = {fw, bw};
  for (int dir = 0; dir < 2; ++dir) {
    a.kernel[dir] = P[dir]->kernel; a.wi[dir] = P[dir]->w_i_diag; a.wf[dir] = P[dir]->w_f_diag; a.wo[dir] = P[dir]->w_o_diag;
  }
  a.seq_len = seq_len; a.G = G; a.y = y; a.gates = gates; a.cs = cs; a.hs_lp = hs_lp; a.y_lp = y_lp; a.final_state = final_state;
  a.hx = (__nv_bfloat16*)workspace;
  const size_t hx_bytes = align_up((size_t)2 * 2 * WB * 4 * d->H * 2, 256);
  a.bar = (unsigned*)((char*)workspace + hx_bytes);
  B2_CUDA(cudaMemsetAsync(a.bar, 0, 2 * 64 * sizeof(unsigned), stream));
#if B2_REC_TIMING
  if (env_int("B2_REC_DBG", 0)) {
    int rcl = B2_ERR_UNSUPPORTED;
    if (d->H / 32 == 32) rcl = launch_wide_fwd<32>(a, stream);
    long long hb[16];
    cudaMemcpyAsync(hb, (char*)a.bar + 512, sizeof(hb), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    for (int w = 0; w < 2; ++w)
      fprintf(stderr, "[wide fwd dbg warp %d] cycles/step: flags=%lld until_h=%lld mma=%lld part+sync+reduce=%lld gate+stores=%lld "
              "publish_sync=%lld fence+flag=%lld | step=%lld\n", w * 4, hb[w * 8 + 0] / a.T, hb[w * 8 + 1] / a.T, hb[w * 8 + 2] / a.T,
              hb[w * 8 + 3] / a.T, hb[w * 8 + 4] / a.T, hb[w * 8 + 5] / a.T, hb[w * 8 + 6] / a.T, hb[w * 8 + 7] / a.T);
    return rcl;
  }
#endif
  switch (d->H / 32) {
    case 20: return launch_wide_fwd<20>(a, stream);
    case 24: return launch_wide_fwd<24>(a, stream);
    case 28: return launch_wide_fwd<28>(a, stream);
    case 32: return launch_wide_fwd<32>(a, stream);
  }
  set_error("wide_rec_forward: unsupported H=%d", d->H);
  return B2_ERR_UNSUPPORTED;
}

template <int KS8>
static int launch_wide_bwd(WideBwdArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)3 * WB * (a.H + 8) * 2 + (size_t)8 * 16 * WPP * 4 + 64;
  auto kern = lstm_wide_bwd_kernel<KS8>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {(void*)&a};
  B2_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(2 * (a.H / WU)), dim3(WTHREADS), params, smem, stream));
  count_launches(1);
  return B2_OK;
}

// BPTT recurrence of a wide layer -> dG [T*B, 8H] fp32 (TF column order); same workspace block as the forward pass
int wide_rec_backward(const b2_lstm_desc* d, const b2_lstm_params* fw, const b2_lstm_params* bw, const int32_t* seq_len,
                      const float* dy, const float* gates, const float* cs, const float* d_final_state, float* dG,
                      __nv_bfloat16* dG_lp, void* workspace, cudaStream_t stream) {
  WideBwdArgs a;
  a.T = d->T; a.B = d->B; a.D_in = d->D_in; a.H = d->H;
  a.use_peephole = d->use_peephole; a.cell_clip = d->cell_clip; a.keep_prob = d->keep_prob; a.seed = d->dropout_seed;
  const b2_lstm_params* P[2] = {fw, bw};
  for (int dir = 0; dir < 2; ++dir) {
    a.kernel[dir] = P[dir]->kernel; a.wi[dir] = P[dir]->w_i_diag; a.wf[dir] = P[dir]->w_f_diag; a.wo[dir] = P[dir]->w_o_diag;
  }
  a.seq_len = seq_len; a.dy = dy; a.gates = gates; a.cs = cs; a.dfinal = d_final_state; a.dG = dG; a.dG_lp = dG_lp;
  a.dzx = (__nv_bfloat16*)workspace;
  const size_t zx_bytes = align_up((size_t)2 * 2 * WB * 4 * d->H * 2, 256);
  a.bar = (unsigned*)((char*)workspace + zx_bytes);
  B2_CUDA(cudaMemsetAsync(a.bar, 0, 2 * 64 * sizeof(unsigned), stream));
  switch (d->H / 128) {
    case 5: return launch_wide_bwd<5>(a, stream);
    case 6: return launch_wide_bwd<6>(a, stream);
    case 7: return launch_wide_bwd<7>(a, stream);
    case 8: return launch_wide_bwd<8>(a, stream);
  }
  set_error("wide_rec_backward: unsupported H=%d", d->H);
  return B2_ERR_UNSUPPORTED;
}

}  // namespace b2
