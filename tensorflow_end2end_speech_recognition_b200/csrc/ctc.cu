// CTC forward-backward loss + gradient for sm_100a.
//
// Replaces tf.nn.ctc_loss as called at models/ctc/ctc.py:289-297 (reference
// arithmetic lives in TensorFlow; semantics restated in oracle/ctc.py).
//
// Kernels, all HBM/latency bound (no GEMM shape anywhere):
//   short labels (2*Lmax+1 <= 512, every BASELINE config), the fast path:
//     ctc_ab_team_kernel<SPL>  four warps per (utterance, direction): a thread keeps SPL consecutive lattice
//                              positions in registers, neighbours come by shuffle (two shared cells across a warp
//                              boundary) -- no shared-memory column, one block barrier per lattice step; the
//                              lattice runs on RAW logits (the per-frame normaliser is added by the finaliser),
//                              emissions prefetched by a cp.async ring; rows spilled [B,T,S_pad] lane-interleaved
//     ctc_softmax_rows_kernel  one warp per (t,b) row, the row in registers (all of its loads in flight at once):
//                              logsumexp (written to lse[t,b]), softmax streamed out; on a second stream, BESIDE the
//                              sweep (it needs no lattice)
//     ctc_occ_rows_kernel      occupancies (row-local normalisation) subtracted from the stored rows with red.global
//                              (C > 3072: ctc_grad_kernel, the row in shared memory, does both after the sweep)
//     ctc_finalize_kernel      loss[b] = -(log p' - sum_t lse[t,b])
//   long labels (fallback): ctc_lse_kernel, ctc_alpha_beta_kernel (one CTA per (utterance, direction), column in
//     shared memory), ctc_grad_kernel reading the precomputed lse
//
// Algorithmic HBM bytes: 8*T*B*C (read logits, write grad) + 16*T*B*S spill.
#include "common.cuh"
#include <stdlib.h>

namespace b2 {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ctc_lse_kernel(const float* __restrict__ logits, const int* __restrict__ seq_len,
               int T, int B, int C, float* __restrict__ lse) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (row >= (int64_t)T * B) return;
  const int t = (int)(row / B), b = (int)(row % B);
  if (t >= seq_len[b]) {
    if (lane == 0) lse[row] = 0.f;
    return;
  }
  const float* x = logits + row * C;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(x[c] - m);
  s = warp_sum(s);
  if (lane == 0) lse[row] = m + __logf(s);
}

// One CTA = one utterance x one direction (blockIdx.y: 0 alpha, 1 beta).
// Thread tid owns lattice positions s = tid + k*blockDim.x, k < SPT.
template <int SPT>
__global__ void __launch_bounds__(1024)
ctc_alpha_beta_kernel(const float* __restrict__ logits, const float* __restrict__ lse,
                      const int* __restrict__ labels_flat, const int* __restrict__ label_offsets,
                      const int* __restrict__ seq_len, int T, int B, int C, int blank,
                      int S_pad, int ignore_longer, float* __restrict__ alpha,
                      float* __restrict__ beta, float* __restrict__ logp_out,
                      int* __restrict__ skip_out, float* __restrict__ loss) {
  extern __shared__ float smem[];
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int NT = blockDim.x, tid = threadIdx.x;
  const int Tb = min(seq_len[b], T);
  const int off = label_offsets[b];
  const int L = label_offsets[b + 1] - off;
  const int S = 2 * L + 1;
  const int* lab = labels_flat + off;

  if (L > Tb && ignore_longer) {   // skipped utterance: loss 0, grad 0 (ctc.py:296)
    if (tid == 0 && !is_beta) { loss[b] = 0.f; logp_out[b] = 0.f; skip_out[b] = 1; }
    return;
  }
  if (Tb <= 0) {
    if (tid == 0 && !is_beta) {
      loss[b] = (L == 0) ? 0.f : INFINITY;
      logp_out[b] = (L == 0) ? 0.f : -INFINITY;
      skip_out[b] = 1;
    }
    return;
  }
  if (tid == 0 && !is_beta) skip_out[b] = 0;

  const int W = S_pad + 4;          // own cell of position s is buf[s + 2]
  float* buf0 = smem;
  float* buf1 = smem + W;
  for (int i = tid; i < 2 * W; i += NT) smem[i] = -INFINITY;

  int cls[SPT];
  bool skip[SPT], valid[SPT];
#pragma unroll
  for (int k = 0; k < SPT; ++k) {
    const int s = tid + k * NT;
    valid[k] = s < S;
    cls[k] = (valid[k] && (s & 1)) ? lab[s >> 1] : blank;
    skip[k] = false;
    if (valid[k] && (s & 1)) {
      if (!is_beta) skip[k] = (s >= 3) && (lab[s >> 1] != lab[(s >> 1) - 1]);
      else          skip[k] = (s + 2 < S) && (lab[(s >> 1) + 1] != lab[s >> 1]);
    }
  }
  // label values index logits rows and shared gradient bins: reject anything outside [0, C) or equal to
  // the blank (tf.nn.ctc_loss raises InvalidArgument; here the utterance gets loss NaN and a zero gradient,
  // the host wrapper raises before launching)
  {
    bool bad = false;
#pragma unroll
    for (int k = 0; k < SPT; ++k)
      if (valid[k] && ((tid + k * NT) & 1)) bad |= (cls[k] < 0) | (cls[k] >= C) | (cls[k] == blank);
    if (__syncthreads_or(bad)) {
      if (tid == 0 && !is_beta) { loss[b] = __int_as_float(0x7fc00000); logp_out[b] = 0.f; skip_out[b] = 1; }
      return;
    }
  }
  float* out = (is_beta ? beta : alpha) + (int64_t)b * T * S_pad;
  const int t0 = is_beta ? Tb - 1 : 0;
  const int dt = is_beta ? -1 : 1;

  // t = t0 (initial column)
  {
    const int64_t row = (int64_t)t0 * B + b;
    const float l = lse[row];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int s = tid + k * NT;
      if (!valid[k]) continue;
      float v = -INFINITY;
      const bool init = is_beta ? (s >= S - 2) : (s <= 1);
      if (init) v = logits[row * C + cls[k]] - l;
      buf0[s + 2] = v;
      out[(int64_t)t0 * S_pad + s] = v;
    }
  }
  __syncthreads();

  // Every kNorm steps the column is shifted by its maximum so that lattice values stay
  // O(100) in magnitude (fp32 ulp ~1e-5) instead of O(T).  Rows spilled to HBM therefore
  // carry arbitrary per-row offsets; the gradient kernel normalises each row locally
  // (sum_s alpha*beta/y is the same constant p for every t), the loss adds the offsets back.
  constexpr int kNorm = 16;
  __shared__ float s_red[32];
  __shared__ float s_shift;
  double offset_sum = 0.0;           // meaningful in thread 0 only

  // prefetch ring for the emission gather.  A register ring fed by ld.global does not work: the
  // compiler tracks loads with 6 scoreboards, the refill issued at step i shares one with the slot
  // consumed at step i+1, so every step waited a full memory latency (ncu: 28 % long-scoreboard on
  // the FADD that consumes the slot, ~1000 cycles per lattice step).  cp.async groups are counted,
  // not scoreboarded: the ring lives in shared memory, every thread copies and later reads only
  // its own cells, `cp.async.wait_group PF-1` releases exactly the oldest slot.
  constexpr int PF = 4;
  float* ring = smem + 2 * W;                       // [PF][SPT*NT + 32]  (+32: the row's lse, one copy per warp)
  const int ring_stride = SPT * NT + 32;
  const int nsteps = Tb - 1;
  // running pointers (one 64-bit add per step instead of re-deriving row*C + class every time)
  const int64_t row_step = (int64_t)dt * B;
  const float* lg_next[SPT];
#pragma unroll
  for (int k = 0; k < SPT; ++k) lg_next[k] = logits + ((int64_t)(t0 + dt) * B + b) * C + cls[k];
  const float* lse_next = lse + (int64_t)(t0 + dt) * B + b;
  const int64_t lg_stride = row_step * C;
  int issued = 0;
  auto issue = [&](int j) {                         // emissions of the next not-yet-issued lattice step -> slot j
    if (issued < nsteps) {
      float* slot = ring + (size_t)j * ring_stride;
#pragma unroll
      for (int k = 0; k < SPT; ++k) {
        cp_async4(slot + k * NT + tid, lg_next[k]);
        lg_next[k] += lg_stride;
      }
      if ((tid & 31) == 0) cp_async4(slot + SPT * NT + (tid >> 5), lse_next);
      lse_next += row_step;
    }
    ++issued;
    cp_async_commit();                              // empty groups keep the count uniform
  };
#pragma unroll
  for (int j = 0; j < PF; ++j) issue(j);
  float* out_next = out + (int64_t)(t0 + dt) * S_pad + tid;
  const int64_t out_stride = (int64_t)dt * S_pad;
  float* prev = buf0;
  float* cur = buf1;
  for (int i0 = 0; i0 < nsteps; i0 += PF) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int i = i0 + j;
      if (i < nsteps) {               // uniform across the CTA
        cp_async_wait<PF - 1>();      // slot j has landed (this thread's own copies)
        __syncwarp();                 // ... and lane 0's copy of the row's lse is visible to its warp
        const float* slot = ring + (size_t)j * ring_stride;
        const float lq = slot[SPT * NT + (tid >> 5)];     // copied by lane 0 of this warp
        float xv[SPT];
#pragma unroll
        for (int k = 0; k < SPT; ++k) xv[k] = slot[k * NT + tid];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const int s = tid + k * NT;
          if (valid[k]) {
            const float a0 = prev[s + 2];
            const float a1 = is_beta ? prev[s + 3] : prev[s + 1];
            const float a2 = skip[k] ? (is_beta ? prev[s + 4] : prev[s]) : -INFINITY;
            const float v = lse3_nb(a0, a1, a2) + (xv[k] - lq);
            cur[s + 2] = v;
            out_next[k * NT] = v;
          }
        }
        out_next += out_stride;
        // refill this ring slot with the emission PF steps ahead
        issue(j);
        __syncthreads();
        float* tmp = prev; prev = cur; cur = tmp;
        if ((i % kNorm) == kNorm - 1) {     // uniform: renormalise the column just written
          float m = -INFINITY;
#pragma unroll
          for (int k = 0; k < SPT; ++k)
            if (valid[k]) m = fmaxf(m, prev[tid + k * NT + 2]);
          m = warp_max(m);
          if ((tid & 31) == 0) s_red[tid >> 5] = m;
          __syncthreads();
          if (tid < 32) {
            float mm = (tid < (NT + 31) / 32) ? s_red[tid] : -INFINITY;
            mm = warp_max(mm);
            if (tid == 0) s_shift = mm;
          }
          __syncthreads();
          const float sh = s_shift;
          if (sh != -INFINITY) {
#pragma unroll
            for (int k = 0; k < SPT; ++k)
              if (valid[k]) prev[tid + k * NT + 2] -= sh;
            offset_sum += (double)sh;
          }
          __syncthreads();
        }
      }
    }
  }
  if (!is_beta && tid == 0) {
    const float a = prev[(S - 1) + 2];
    const float c = (S > 1) ? prev[(S - 2) + 2] : -INFINITY;
    const float lp = (float)((double)lse2(a, c) + offset_sum);
    logp_out[b] = lp;
    loss[b] = -lp;
  }
}

// base-2 log-domain helpers of the team sweep: ex2/lg2 are the hardware functions, the .ftz forms carry no
// denormal range fix-ups (3 instructions per exp, 3 per log in the default forms) -- a term 2^-126 below the column
// maximum contributes nothing at fp32 anyway
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float l2se3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, fmaxf(b, c)), -1e30f);
  return m + lg2_ftz(ex2_ftz(a - m) + ex2_ftz(b - m) + ex2_ftz(c - m));
}
__device__ __forceinline__ float l2se2(float a, float b) {
  const float m = fmaxf(fmaxf(a, b), -1e30f);
  return m + lg2_ftz(ex2_ftz(a - m) + ex2_ftz(b - m));
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---------------------------------------------------------------- fast path: four warps per (utterance, direction)
// Thread `tid` (of kTeam = 128) owns positions s = tid*SPL + k (k < SPL, SPL even: k odd <=> label position) in
// registers.  Neighbours inside a warp come by shuffle, the two values across a warp boundary through a parity-
// double-buffered shared cell pair -> ONE block barrier per lattice step, no shared-memory column.  (One warp with
// 16 positions per lane was measured first: 3 000 cycles per step, issue-bound on a single scheduler.)
// The lattice is kept in BASE-2 log units (values = log2 of the unnormalised path mass); the spilled rows too.
// Spill index of position s inside a row: k*kTeam + tid (coalesced); S_pad = kTeam*SPL.
constexpr int kTeam = 128;
template <int SPL, bool is_beta>
__device__ __forceinline__ void
ctc_ab_team_body(const float* __restrict__ logits, const int* __restrict__ labels_flat,
                 const int* __restrict__ label_offsets, const int* __restrict__ seq_len, int T, int B, int C,
                 int blank, int ignore_longer, float* __restrict__ alpha, float* __restrict__ beta,
                 float* __restrict__ logp_out, int* __restrict__ skip_out, float* __restrict__ loss) {
  constexpr int S_pad = kTeam * SPL;
  constexpr int NL = SPL / 2;                 // label positions per thread
  constexpr int NW = kTeam / 32;
  constexpr int PF = 8;                       // emission prefetch depth (lattice steps)
  __shared__ float ring[PF][NL * kTeam + NW]; // [slot][k/2][tid], then the blank's logit (one copy per warp)
  __shared__ float xch[2][NW][2];             // boundary values of the column of parity p
  __shared__ float red[NW];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tb = min(seq_len[b], T);
  const int off = label_offsets[b];
  const int L = label_offsets[b + 1] - off;
  const int S = 2 * L + 1;
  const int* lab = labels_flat + off;

  if (L > Tb && ignore_longer) {   // skipped utterance: loss 0, grad 0 (ctc.py:296)
    if (tid == 0 && !is_beta) { loss[b] = 0.f; logp_out[b] = 0.f; skip_out[b] = 1; }
    return;
  }
  if (Tb <= 0) {
    if (tid == 0 && !is_beta) {
      loss[b] = (L == 0) ? 0.f : INFINITY;
      logp_out[b] = (L == 0) ? 0.f : -INFINITY;
      skip_out[b] = 1;
    }
    return;
  }
  int cls[NL];
  bool skip[NL], valid[SPL];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < SPL; ++k) valid[k] = tid * SPL + k < S;
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    const int s = tid * SPL + 2 * j + 1;
    cls[j] = blank; skip[j] = false;
    if (s < S) {
      cls[j] = lab[s >> 1];
      bad |= (cls[j] < 0) | (cls[j] >= C) | (cls[j] == blank);
      if (!is_beta) skip[j] = (s >= 3) && (lab[s >> 1] != lab[(s >> 1) - 1]);
      else          skip[j] = (s + 2 < S) && (lab[(s >> 1) + 1] != lab[s >> 1]);
    }
  }
  if (__syncthreads_or(bad)) {             // see ctc_alpha_beta_kernel: loss NaN, zero gradient
    if (tid == 0 && !is_beta) { loss[b] = __int_as_float(0x7fc00000); logp_out[b] = 0.f; skip_out[b] = 1; }
    return;
  }
  if (tid == 0 && !is_beta) skip_out[b] = 0;

  float* out = (is_beta ? beta : alpha) + (int64_t)b * T * S_pad;
  const int t0 = is_beta ? Tb - 1 : 0;
  const int dt = is_beta ? -1 : 1;
  float a[SPL];
  {
    const float* x = logits + ((int64_t)t0 * B + b) * C;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
      const int s = tid * SPL + k;
      const bool init = is_beta ? (s >= S - 2) : (s <= 1);
      a[k] = -INFINITY;
      if (valid[k] && init) a[k] = kLog2e * ((k & 1) ? x[cls[k >> 1]] : x[blank]);
      out[(int64_t)t0 * S_pad + k * kTeam + tid] = a[k];
    }
  }
  // publish the boundary values of column 0 (parity 0)
  if (!is_beta) { if (lane == 31) { xch[0][warp][0] = a[SPL - 2]; xch[0][warp][1] = a[SPL - 1]; } }
  else          { if (lane == 0)  { xch[0][warp][0] = a[0];       xch[0][warp][1] = a[1]; } }
  const int nsteps = Tb - 1;
  const int64_t lg_stride = (int64_t)dt * B * C;
  const float* lg_next[NL];
#pragma unroll
  for (int j = 0; j < NL; ++j) lg_next[j] = logits + ((int64_t)(t0 + dt) * B + b) * C + cls[j];
  const float* blank_next = logits + ((int64_t)(t0 + dt) * B + b) * C + blank;
  int issued = 0;
  auto issue = [&](int slot) {
    if (issued < nsteps) {
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        cp_async4(&ring[slot][j * kTeam + tid], lg_next[j]);
        lg_next[j] += lg_stride;
      }
      if (lane == 0) cp_async4(&ring[slot][NL * kTeam + warp], blank_next);
      blank_next += lg_stride;
    }
    ++issued;
    cp_async_commit();
  };
#pragma unroll
  for (int j = 0; j < PF; ++j) issue(j);
  float* out_next = out + (int64_t)(t0 + dt) * S_pad + tid;
  const int64_t out_stride = (int64_t)dt * S_pad;
  constexpr int kNorm = 16;
  double offset_sum = 0.0;
  __syncthreads();
  for (int i0 = 0; i0 < nsteps; i0 += PF) {
#pragma unroll
    for (int jj = 0; jj < PF; ++jj) {
      const int i = i0 + jj;
      if (i < nsteps) {                 // uniform across the CTA
        const int par = i & 1;          // parity of the column being read
        cp_async_wait<PF - 2>();        // emissions of this step have landed (the refill runs one step late)
        __syncwarp();
        const float xb = ring[jj][NL * kTeam + warp];
        float xl[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) xl[j] = ring[jj][j * kTeam + tid];
        // neighbours across the thread boundary
        float e0, e1;       // alpha: positions s-2, s-1 of this thread's first position; beta: s+1, s+2 of its last
        if (!is_beta) {
          e0 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
          e1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);
          if (lane == 0) {
            e0 = warp > 0 ? xch[par][warp - 1][0] : -INFINITY;
            e1 = warp > 0 ? xch[par][warp - 1][1] : -INFINITY;
          }
        } else {
          e0 = __shfl_down_sync(0xffffffffu, a[0], 1);
          e1 = __shfl_down_sync(0xffffffffu, a[1], 1);
          if (lane == 31) {
            e0 = warp < NW - 1 ? xch[par][warp + 1][0] : -INFINITY;
            e1 = warp < NW - 1 ? xch[par][warp + 1][1] : -INFINITY;
          }
        }
        // refill the slot consumed in the PREVIOUS step (every lane passed that step's block barrier since it
        // read the slot): independent instructions that fill the latency holes of the chain below
        if (i > 0) issue((jj + PF - 1) % PF);
        float v[SPL];
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
          float n1, n2;                        // k is a compile-time constant after unrolling
          if (!is_beta) {
            n1 = (k >= 1) ? a[(k + SPL - 1) % SPL] : e1;                     // s-1
            n2 = (k >= 2) ? a[(k + SPL - 2) % SPL] : (k == 1 ? e1 : e0);     // s-2
          } else {
            n1 = (k + 1 < SPL) ? a[(k + 1) % SPL] : e0;                      // s+1
            n2 = (k + 2 < SPL) ? a[(k + 2) % SPL] : (k + 2 == SPL ? e0 : e1);   // s+2
          }
          float r;
          if (k & 1) r = fmaf(xl[k >> 1], kLog2e, l2se3(a[k], n1, skip[k >> 1] ? n2 : -INFINITY));
          else       r = fmaf(xb, kLog2e, l2se2(a[k], n1));            // blanks have no skip transition
          v[k] = valid[k] ? r : -INFINITY;
        }
        if ((i % kNorm) == kNorm - 1) {       // uniform: shift the new column by its maximum before publishing it
          float m = v[0];
#pragma unroll
          for (int k = 1; k < SPL; ++k) m = fmaxf(m, v[k]);
          m = warp_max(m);
          if (lane == 0) red[warp] = m;
          __syncthreads();
          m = red[0];
#pragma unroll
          for (int w = 1; w < NW; ++w) m = fmaxf(m, red[w]);
          if (m != -INFINITY) {
#pragma unroll
            for (int k = 0; k < SPL; ++k) v[k] -= m;
            offset_sum += (double)m;
          }
        }
#pragma unroll
        for (int k = 0; k < SPL; ++k) { a[k] = v[k]; out_next[k * kTeam] = v[k]; }
        out_next += out_stride;
        if (!is_beta) { if (lane == 31) { xch[par ^ 1][warp][0] = a[SPL - 2]; xch[par ^ 1][warp][1] = a[SPL - 1]; } }
        else          { if (lane == 0)  { xch[par ^ 1][warp][0] = a[0];       xch[par ^ 1][warp][1] = a[1]; } }
        __syncthreads();
      }
    }
  }
  if (!is_beta) {
    float e = -INFINITY;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
      const int s = tid * SPL + k;
      if (s == S - 1 || s == S - 2) e = lse2(e, a[k] * kLn2);
    }
    // at most two threads hold a finite e (natural log from here on)
    float m = warp_max(e);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float z = (m != -INFINITY) ? __expf(e - m) : 0.f;
    z = warp_sum(z);
    if (lane == 0) red[warp] = z;
    __syncthreads();
    if (tid == 0) {
      float zz = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) zz += red[w];
      const float lp = (m != -INFINITY) ? m + __logf(zz) : -INFINITY;
      logp_out[b] = (float)((double)lp + offset_sum * 0.6931471805599453);     // still lacks - sum_t lse[t,b]
    }
  }
}

template <int SPL>
__global__ void __launch_bounds__(kTeam)
ctc_ab_team_kernel(const float* __restrict__ logits, const int* __restrict__ labels_flat,
                   const int* __restrict__ label_offsets, const int* __restrict__ seq_len, int T, int B, int C,
                   int blank, int ignore_longer, float* __restrict__ alpha, float* __restrict__ beta,
                   float* __restrict__ logp_out, int* __restrict__ skip_out, float* __restrict__ loss) {
  // blockIdx.y: 0 alpha, 1 beta -- two instantiations of the body, no per-step direction branches
  if (blockIdx.y == 0)
    ctc_ab_team_body<SPL, false>(logits, labels_flat, label_offsets, seq_len, T, B, C, blank, ignore_longer, alpha, beta,
                                 logp_out, skip_out, loss);
  else
    ctc_ab_team_body<SPL, true>(logits, labels_flat, label_offsets, seq_len, T, B, C, blank, ignore_longer, alpha, beta,
                                logp_out, skip_out, loss);
}

// loss[b] = -(log p' - sum_{t < T_b} lse[t,b]) for the utterances the sweep did not settle itself
__global__ void __launch_bounds__(128)
ctc_finalize_kernel(const float* __restrict__ lse, const int* __restrict__ seq_len,
                    const int* __restrict__ skip, int T, int B, float* __restrict__ logp,
                    float* __restrict__ loss) {
  __shared__ double part[4];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (skip[b]) return;
  const int Tb = min(seq_len[b], T);
  double acc = 0.0;
  for (int t = tid; t < Tb; t += 128) acc += (double)__ldg(lse + (int64_t)t * B + b);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    const float lp = (float)((double)logp[b] - (part[0] + part[1] + part[2] + part[3]));
    logp[b] = lp;
    loss[b] = -lp;
  }
}

// One warp per (t,b) row.  g[c] = softmax_c - sum_{s: l'(s)=c} alpha*beta/(y*Z_t) with the
// row-local normaliser Z_t = sum_s alpha*beta/y (= p, independent of per-row lattice shifts and of whether the lattice
// ran on normalised or raw logits).  spl > 0: rows of the team sweep (position tid*spl + k at k*nt + tid, lattice on
// raw logits, this kernel computes and stores the row's logsumexp); spl == 0: linear rows, lse precomputed.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ctc_grad_kernel(const float* __restrict__ logits, float* __restrict__ lse,
                const int* __restrict__ labels_flat, const int* __restrict__ label_offsets,
                const int* __restrict__ seq_len, const float* __restrict__ alpha,
                const float* __restrict__ beta, const int* __restrict__ skip_in, int T, int B, int C, int blank,
                int S_pad, int spl, int nt, float grad_scale, int warps_per_block, float* __restrict__ grad) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * warps_per_block + warp;
  if (row >= (int64_t)T * B) return;
  const int t = (int)(row / B), b = (int)(row % B);
  float* g = smem + (size_t)warp * C;
  float* out = grad + row * C;
  const int Tb = min(seq_len[b], T);
  if (t >= Tb || skip_in[b]) {
    if (grad)
      for (int c = lane; c < C; c += 32) out[c] = 0.f;
    if (spl > 0 && lane == 0) lse[row] = 0.f;
    return;
  }
  const float* x = logits + row * C;
  float l, unit = 1.f;         // g[] holds softmax * unit
  if (spl > 0) {
    // the row once through registers (kLd independent 128-byte loads per warp in flight: at C = 3001 the kernel is
    // HBM-bound and 8 in flight measured 2.2 TB/s) into shared memory, running maximum
    constexpr int kLd = 24;
    float m = -INFINITY;
    for (int c0 = 0; c0 < C; c0 += 32 * kLd) {
      float v[kLd];
#pragma unroll
      for (int j = 0; j < kLd; ++j) {
        const int c = c0 + j * 32 + lane;
        v[j] = c < C ? __ldg(x + c) : -INFINITY;
      }
#pragma unroll
      for (int j = 0; j < kLd; ++j) {
        const int c = c0 + j * 32 + lane;
        if (c < C) { g[c] = v[j]; m = fmaxf(m, v[j]); }
      }
    }
    m = warp_max(m);
    __syncwarp();
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) { const float e = __expf(g[c] - m); g[c] = e; sum += e; }
    sum = warp_sum(sum);
    l = m + __logf(sum);
    if (lane == 0) lse[row] = l;
    if (!grad) return;
    unit = sum;                // normalised in the output pass
  } else {
    l = lse[row];
    for (int c = lane; c < C; c += 32) g[c] = __expf(x[c] - l);
  }
  __syncwarp();
  const float out_scale = grad_scale / unit;
  {
    const int off = label_offsets[b];
    const int L = label_offsets[b + 1] - off;
    const int S = 2 * L + 1;
    const float* ar = alpha + ((int64_t)b * T + t) * S_pad;
    const float* br = beta + ((int64_t)b * T + t) * S_pad;
    if (spl > 0) {
      // team-sweep rows (raw-logit lattice, <= 512 stored cells): every lane keeps its <= 16 terms
      // alpha + beta - x[class] and their classes in registers -- one read of the lattice rows, one label gather
      constexpr int kMax = 16;
      const int n_idx = nt * spl;
      float d[kMax]; int cc[kMax];
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < kMax; ++j) {
        const int i = j * 32 + lane;
        d[j] = -INFINITY; cc[j] = -1;
        if (i < n_idx) {
          const int s = (i & (kTeam - 1)) * spl + (i / kTeam);       // nt == kTeam
          if (s < S) {
            cc[j] = (s & 1) ? labels_flat[off + (s >> 1)] : blank;
            d[j] = fmaf(ar[i] + br[i], kLn2, -__ldg(x + cc[j]));     // rows are in base-2 log units
            m = fmaxf(m, d[j]);
          }
        }
      }
      m = warp_max(m);
      if (m != -INFINITY) {        // -inf: no alignment passes through this frame -> grad = softmax
        float z = 0.f;
#pragma unroll
        for (int j = 0; j < kMax; ++j) { d[j] = __expf(d[j] - m); z += d[j]; }     // exp(-inf) = 0 for the unused cells
        z = warp_sum(z);
        const float k = unit / z;
        float blank_occ = 0.f;
#pragma unroll
        for (int j = 0; j < kMax; ++j) {
          if (cc[j] == blank) blank_occ += d[j];
          else if (cc[j] >= 0) atomicAdd(&g[cc[j]], -d[j] * k);
        }
        blank_occ = warp_sum(blank_occ);
        __syncwarp();
        if (lane == 0) g[blank] -= blank_occ * k;
        __syncwarp();
      }
    } else {
      const float xb = x[blank] - l;
      // pass 1: Z = logsumexp_s (alpha + beta - lp): equals log p up to the per-row shifts
      float m = -INFINITY;
      for (int s = lane; s < S; s += 32) {
        const float lp = (s & 1) ? (x[labels_flat[off + (s >> 1)]] - l) : xb;
        m = fmaxf(m, ar[s] + br[s] - lp);
      }
      m = warp_max(m);
      if (m != -INFINITY) {
        float z = 0.f;
        for (int s = lane; s < S; s += 32) {
          const float lp = (s & 1) ? (x[labels_flat[off + (s >> 1)]] - l) : xb;
          z += __expf(ar[s] + br[s] - lp - m);
        }
        z = warp_sum(z);
        const float Z = m + __logf(z);
        float blank_occ = 0.f;
        for (int s = 2 * lane; s < S; s += 64) blank_occ += __expf(ar[s] + br[s] - xb - Z);
        for (int s = 2 * lane + 1; s < S; s += 64) {
          const int c = labels_flat[off + (s >> 1)];
          atomicAdd(&g[c], -__expf(ar[s] + br[s] - (x[c] - l) - Z));
        }
        blank_occ = warp_sum(blank_occ);
        __syncwarp();
        if (lane == 0) g[blank] -= blank_occ;
        __syncwarp();
      }
    }
  }
  for (int c = lane; c < C; c += 32) out[c] = out_scale * g[c];
}

// Row passes of the team-sweep path for C <= 32 * NV (NV <= 96).
// (1) ctc_softmax_rows_kernel: needs no lattice -> runs on a second stream BESIDE the sweep.  The whole row lives in
//     REGISTERS (every load of the row in flight at once: 12 KB per warp at C = 3001 -- a shared-memory version with
//     24 x 128 B in flight measured 2.9 TB/s): logsumexp -> lse[t,b], grad_scale * softmax streamed straight out.
// (2) ctc_occ_rows_kernel, after both: the <= 512 lattice terms of the row cached in registers, occupancies (row-local
//     normalisation) subtracted from the stored row with red.global.
template <int NV>
__global__ void __launch_bounds__(128)
ctc_softmax_rows_kernel(const float* __restrict__ logits, float* __restrict__ lse, const int* __restrict__ seq_len,
                        int T, int B, int C, float grad_scale, float* __restrict__ grad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + warp;
  if (row >= (int64_t)T * B) return;
  const int t = (int)(row / B), b = (int)(row % B);
  float* out = grad + row * C;
  if (t >= min(seq_len[b], T)) {
    if (grad)
      for (int c = lane; c < C; c += 32) out[c] = 0.f;
    if (lane == 0) lse[row] = 0.f;
    return;
  }
  const float* x = logits + row * C;
  float v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = j * 32 + lane;
    v[j] = c < C ? __ldg(x + c) : -INFINITY;
  }
  float m = v[0];
#pragma unroll
  for (int j = 1; j < NV; ++j) m = fmaxf(m, v[j]);
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) { v[j] = __expf(v[j] - m); sum += v[j]; }
  sum = warp_sum(sum);
  if (lane == 0) lse[row] = m + __logf(sum);
  if (!grad) return;
  const float k0 = grad_scale / sum;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = j * 32 + lane;
    if (c < C) out[c] = v[j] * k0;
  }
}

__global__ void __launch_bounds__(128)
ctc_occ_rows_kernel(const float* __restrict__ logits, const int* __restrict__ labels_flat,
                    const int* __restrict__ label_offsets, const int* __restrict__ seq_len,
                    const float* __restrict__ alpha, const float* __restrict__ beta, const int* __restrict__ skip_in,
                    int T, int B, int C, int blank, int S_pad, int spl, float grad_scale, float* __restrict__ grad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + warp;
  if (row >= (int64_t)T * B) return;
  const int t = (int)(row / B), b = (int)(row % B);
  float* out = grad + row * C;
  if (t >= min(seq_len[b], T)) return;
  if (skip_in[b]) {                  // skipped / rejected utterance: zero gradient (the softmax pass could not know)
    for (int c = lane; c < C; c += 32) out[c] = 0.f;
    return;
  }
  const float* x = logits + row * C;
  const int off = label_offsets[b];
  const int L = label_offsets[b + 1] - off;
  const int S = 2 * L + 1;
  const float* ar = alpha + ((int64_t)b * T + t) * S_pad;
  const float* br = beta + ((int64_t)b * T + t) * S_pad;
  constexpr int kMax = 16;
  const int n_idx = kTeam * spl;
  float d[kMax]; int cc[kMax];
  float dm = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMax; ++j) {
    const int i = j * 32 + lane;
    d[j] = -INFINITY; cc[j] = -1;
    if (i < n_idx) {
      const int s = (i & (kTeam - 1)) * spl + (i / kTeam);
      if (s < S) {
        cc[j] = (s & 1) ? labels_flat[off + (s >> 1)] : blank;
        d[j] = fmaf(ar[i] + br[i], kLn2, -__ldg(x + cc[j]));     // rows are in base-2 log units, raw logits
        dm = fmaxf(dm, d[j]);
      }
    }
  }
  dm = warp_max(dm);
  if (dm == -INFINITY) return;       // no alignment passes through this frame -> grad = softmax
  float z = 0.f;
#pragma unroll
  for (int j = 0; j < kMax; ++j) { d[j] = __expf(d[j] - dm); z += d[j]; }
  z = warp_sum(z);
  const float k1 = -grad_scale / z;
  float blank_occ = 0.f;
#pragma unroll
  for (int j = 0; j < kMax; ++j) {
    if (cc[j] == blank) blank_occ += d[j];
    else if (cc[j] >= 0) atomicAdd(out + cc[j], d[j] * k1);
  }
  blank_occ = warp_sum(blank_occ);
  if (lane == 0) atomicAdd(out + blank, blank_occ * k1);
}

// second stream + fork/join events of the row pass (one set per device)
struct CtcSide { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static CtcSide g_ctc_side[16];
static CtcSide* ctc_side() {
  int dev = 0;
  cudaGetDevice(&dev);
  CtcSide* c = &g_ctc_side[dev & 15];
  if (!c->s) {
    cudaStreamCreateWithFlags(&c->s, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->join, cudaEventDisableTiming);
  }
  return c;
}

struct CtcWs {
  float* lse; float* alpha; float* beta; float* logp; int* skip; int S_pad;
};

static size_t ctc_ws_layout(int T, int B, int max_label_len, void* base, CtcWs* w) {
  const int S_max_ = 2 * max_label_len + 1;
  // team sweep (S <= 512): kTeam * SPL with SPL in {2, 4}; fallback: linear rows
  const int S_pad = S_max_ <= 2 * kTeam ? 2 * kTeam : (S_max_ <= 4 * kTeam ? 4 * kTeam : (int)align_up((size_t)S_max_, 32));
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  size_t o_lse = take((size_t)T * B * 4);
  size_t o_a = take((size_t)B * T * S_pad * 4);
  size_t o_b = take((size_t)B * T * S_pad * 4);
  size_t o_lp = take((size_t)B * 4);
  size_t o_sk = take((size_t)B * 4);
  if (w) {
    char* p = (char*)base;
    w->lse = (float*)(p + o_lse); w->alpha = (float*)(p + o_a); w->beta = (float*)(p + o_b);
    w->logp = (float*)(p + o_lp); w->skip = (int*)(p + o_sk); w->S_pad = S_pad;
  }
  return off;
}

}  // namespace b2

using namespace b2;

extern "C" size_t b2_ctc_workspace_bytes(int T, int B, int C, int max_label_len) {
  (void)C;
  return ctc_ws_layout(T, B, max_label_len, nullptr, nullptr);
}

extern "C" int b2_ctc_loss_grad(const float* logits, const int32_t* labels_flat,
                                const int32_t* label_offsets, const int32_t* seq_len, int T,
                                int B, int C, int blank, int max_label_len,
                                int ignore_longer, float grad_scale, float* loss, float* grad,
                                void* workspace, size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(logits && labels_flat && label_offsets && seq_len && loss && workspace,
               "b2_ctc_loss_grad: null pointer");
  B2_CHECK_ARG(T > 0 && B > 0 && C > 1 && blank >= 0 && blank < C && max_label_len >= 0,
               "b2_ctc_loss_grad: bad shape T=%d B=%d C=%d blank=%d Lmax=%d", T, B, C, blank,
               max_label_len);
  CtcWs w;
  const size_t need = ctc_ws_layout(T, B, max_label_len, workspace, &w);
  if (workspace_bytes < need) {
    set_error("b2_ctc_loss_grad: workspace %zu < %zu", workspace_bytes, need);
    return B2_ERR_WORKSPACE;
  }
  const int64_t rows = (int64_t)T * B;
  const int S_max = 2 * max_label_len + 1;
  const bool warp_path = S_max <= 512;
  int wpb = kWarpsPerBlock;
  while (wpb > 1 && (size_t)wpb * C * 4 > 160 * 1024) wpb >>= 1;
  const size_t gsmem = (size_t)wpb * C * 4;
  B2_CHECK_ARG(gsmem <= 200 * 1024, "b2_ctc_loss_grad: C=%d too large", C);
  B2_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));

  if (warp_path) {
    const int spl = w.S_pad / kTeam;       // 2 or 4
    const int nv = cdiv(C, 32);
    const bool rows_regs = nv <= 96;
    CtcSide* side = rows_regs ? ctc_side() : nullptr;
    if (side) B2_CUDA(cudaEventRecord(side->fork, stream));
    dim3 grid(B, 2);
    if (spl == 2)
      ctc_ab_team_kernel<2><<<grid, kTeam, 0, stream>>>(logits, labels_flat, label_offsets, seq_len, T, B, C, blank,
                                                        ignore_longer, w.alpha, w.beta, w.logp, w.skip, loss);
    else
      ctc_ab_team_kernel<4><<<grid, kTeam, 0, stream>>>(logits, labels_flat, label_offsets, seq_len, T, B, C, blank,
                                                        ignore_longer, w.alpha, w.beta, w.logp, w.skip, loss);
    B2_LAUNCH_CHECK();
    if (rows_regs) {
      // softmax half of the row pass (logsumexp + grad_scale * softmax): no lattice needed -> beside the sweep
      B2_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
#define LAUNCH_ROWS(NV)                                                                                     \
      ctc_softmax_rows_kernel<NV><<<cdiv(rows, 4), 128, 0, side->s>>>(logits, w.lse, seq_len, T, B, C, grad_scale, grad)
      if (nv <= 1) LAUNCH_ROWS(1); else if (nv <= 2) LAUNCH_ROWS(2); else if (nv <= 4) LAUNCH_ROWS(4);
      else if (nv <= 8) LAUNCH_ROWS(8); else if (nv <= 16) LAUNCH_ROWS(16); else if (nv <= 32) LAUNCH_ROWS(32);
      else if (nv <= 64) LAUNCH_ROWS(64); else LAUNCH_ROWS(96);
#undef LAUNCH_ROWS
      B2_LAUNCH_CHECK();
      B2_CUDA(cudaEventRecord(side->join, side->s));
      B2_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
      if (grad) {
        ctc_occ_rows_kernel<<<cdiv(rows, 4), 128, 0, stream>>>(logits, labels_flat, label_offsets, seq_len, w.alpha,
                                                               w.beta, w.skip, T, B, C, blank, w.S_pad, spl, grad_scale,
                                                               grad);
        B2_LAUNCH_CHECK();
      }
    } else {
      ctc_grad_kernel<<<cdiv(rows, wpb), wpb * 32, gsmem, stream>>>(
          logits, w.lse, labels_flat, label_offsets, seq_len, w.alpha, w.beta, w.skip, T, B, C, blank, w.S_pad, spl,
          kTeam, grad_scale, wpb, grad);
      B2_LAUNCH_CHECK();
    }
    ctc_finalize_kernel<<<B, 128, 0, stream>>>(w.lse, seq_len, w.skip, T, B, w.logp, loss);
    B2_LAUNCH_CHECK();
    return B2_OK;
  }

  ctc_lse_kernel<<<cdiv(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream>>>(
      logits, seq_len, T, B, C, w.lse);
  B2_LAUNCH_CHECK();
  {
    int NT = (int)align_up((size_t)S_max, 32);
    int spt = 1;
    while (NT > 1024) { spt *= 2; NT = (int)align_up((size_t)cdiv(S_max, spt), 32); }
    B2_CHECK_ARG(spt <= 8, "b2_ctc_loss_grad: label length %d too long (max 4095)", max_label_len);
    const size_t smem = ((size_t)2 * (w.S_pad + 4) + (size_t)4 * ((size_t)spt * NT + 32)) * sizeof(float);
    dim3 grid(B, 2);
#define LAUNCH_AB(SPT)                                                                      \
    do {                                                                                    \
      B2_CUDA(cudaFuncSetAttribute(ctc_alpha_beta_kernel<SPT>,                              \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      ctc_alpha_beta_kernel<SPT><<<grid, NT, smem, stream>>>(                               \
          logits, w.lse, labels_flat, label_offsets, seq_len, T, B, C, blank, w.S_pad,      \
          ignore_longer, w.alpha, w.beta, w.logp, w.skip, loss);                            \
    } while (0)
    if (spt == 1) LAUNCH_AB(1); else if (spt == 2) LAUNCH_AB(2);
    else if (spt == 4) LAUNCH_AB(4); else LAUNCH_AB(8);
#undef LAUNCH_AB
    B2_LAUNCH_CHECK();
  }
  if (grad) {
    ctc_grad_kernel<<<cdiv(rows, wpb), wpb * 32, gsmem, stream>>>(
        logits, w.lse, labels_flat, label_offsets, seq_len, w.alpha, w.beta, w.skip, T, B, C, blank, w.S_pad, 0,
        0, grad_scale, wpb, grad);
    B2_LAUNCH_CHECK();
  }
  return B2_OK;
}
