// Bidirectional LSTM layer (forward + BPTT) for sm_100a -- host orchestration and
// the fp32 CUDA-core recurrence (B2_PREC_FP32, the parity path).
//
// Replaces tf.contrib.rnn.LSTMBlockCell / LSTMCell / BasicLSTMCell under
// tf.nn.bidirectional_dynamic_rnn(sequence_length=...) as built by
// models/encoders/core/blstm.py:258-332 (cell equations restated from
// models/recurrent/layers/lstm.py:142-183, see oracle/lstm.py).
//
// Structure of one layer (both directions):
//   1. time-batched input projection  G[T*B, 8H] = X . [Wx_fw | Wx_bw] + b      (GEMM)
//   2. recurrence, T sequential steps per direction:
//        z = G[t] + h_{t-1} . Wh ; gates ; c_t ; h_t ; length masking ; dropout
//   backward mirrors it: BPTT recurrence producing dG[T*B, 8H], then the
//   time-batched GEMMs dX = dG . Wx^T, dWx = X^T . dG, dWh = Hprev^T . dG,
//   bias/peephole reductions.
// With B2_PREC_BF16 the time-batched GEMMs run on tcgen05 (gemm_tcgen05.cu) and
// the recurrence on the cluster/TMEM kernel (lstm_rec_tc.cu) when available.
#include "lstm_internal.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace b2 {

struct Work {
  float* G;        // [T*B, 8H] gate pre-activations (fwd) / dG (bwd)
  float* hstate;   // [2 parity][2 dir][B][H]
  float* cstate;   // [2 dir][B][H]   (fwd: c ; bwd: dc)
  float* zrec;     // [2 dir][B][4H]  recurrent pre-activations / dh of one step (wide-H path)
  void* wide;      // exchange buffer + barrier counters of the grid-resident wide-layer recurrence (lstm_wide.cu)
  float* mcur;     // [2 dir][B][H]   projection mode: o*tanh(c) of the frame (fwd) / d(o*tanh(c)) (bwd)
  float* hpnew;    // [2 dir][B][P]   projection mode: projected h of the frame (fwd) / its gradient (bwd)
  float* dhp_all;  // [T*B][2P]       projection mode, backward: d(projected h) of every frame
  __nv_bfloat16* xb;   // bf16 operand copies for the tcgen05 GEMMs
  __nv_bfloat16* wb;
  __nv_bfloat16* gb;
};
static size_t pad8z(size_t x) { return (x + 7) / 8 * 8; }
static size_t work_layout(const b2_lstm_desc* d, void* base, Work* w) {
  const size_t TB = (size_t)d->T * d->B;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t oG = take(TB * 8 * d->H * sizeof(float));
  const size_t Pw = d->num_proj > 0 ? (size_t)d->num_proj : 0;
  const size_t oh = take((size_t)2 * 2 * d->B * (d->H > (int)Pw ? d->H : Pw) * sizeof(float));
  const size_t oc = take((size_t)2 * d->B * d->H * sizeof(float));
  const size_t oz = take((size_t)2 * d->B * 4 * d->H * sizeof(float));
  const size_t om = Pw ? take((size_t)2 * d->B * d->H * sizeof(float)) : 0;
  const size_t ohn = Pw ? take((size_t)2 * d->B * Pw * sizeof(float)) : 0;
  const size_t oda = Pw ? take(TB * 2 * Pw * sizeof(float)) : 0;
  const size_t owide = take(wide_rec_workspace_bytes(d));
  size_t oxb = 0, owb = 0, ogb = 0;
  if (d->precision == B2_PREC_BF16) {
    const size_t din = pad8z((size_t)(d->D_in > 2 * d->H ? d->D_in : 2 * d->H));
    oxb = take(TB * din * 2);                                   // X or Hs as bf16
    owb = take((size_t)(d->D_in + d->H + 64) * 8 * d->H * 2);    // weights (both dirs) as bf16
    ogb = take(TB * 8 * d->H * 2);                               // dG as bf16
  }
  if (w) {
    char* p = (char*)base;
    w->G = (float*)(p + oG); w->hstate = (float*)(p + oh); w->cstate = (float*)(p + oc);
    w->zrec = (float*)(p + oz); w->wide = (void*)(p + owide);
    w->mcur = Pw ? (float*)(p + om) : nullptr; w->hpnew = Pw ? (float*)(p + ohn) : nullptr;
    w->dhp_all = Pw ? (float*)(p + oda) : nullptr;
    w->xb = (__nv_bfloat16*)(p + oxb); w->wb = (__nv_bfloat16*)(p + owb);
    w->gb = (__nv_bfloat16*)(p + ogb);
  }
  return off;
}

// ---------------------------------------------------------------------------
// fp32 recurrence step kernels.  CTA tile = 16 batch x 16 units (x 4 gates),
// 256 threads, thread (b,u).  grid = (H/16, ceil(B/16), 2 directions).
// ---------------------------------------------------------------------------
constexpr int RB = 16, RU = 16, RK = 32;

struct StepArgs {
  int T, B, D_in, H, step;          // step index i: fw works on t=i, bw on t=T-1-i
  int use_peephole; float forget_bias, cell_clip, keep_prob; unsigned long long seed;
  const float* kernel[2]; const float* wi[2]; const float* wf[2]; const float* wo[2];
  const int* seq_len;
  const float* G;                   // [T,B,8H]
  float* hstate; float* cstate;
  float* y;                         // [T,B,2H]
  float* gates; float* cs; float* hs;   // reserve (may be null when !need_backward)
  const float* zrec;                // [2,B,4H] h_prev . Wh computed by the caller (wide-H path) or null
  float* mout;                      // [2,B,H] projection mode: o*tanh(c) of this frame (y / h state are written
                                    // by proj_finalize_kernel after the projection GEMM); null otherwise
};

__global__ void __launch_bounds__(256)
lstm_fwd_step_kernel(const StepArgs a) {
  __shared__ float Ws[RK][4 * RU + 1];
  __shared__ float Hs[RB][RK + 1];
  const int dir = blockIdx.z;
  const int u0 = blockIdx.x * RU, b0 = blockIdx.y * RB;
  const int tu = threadIdx.x & 15, tb = threadIdx.x >> 4;
  const int u = u0 + tu, b = b0 + tb;
  const int H = a.H, B = a.B;
  const int t = dir == 0 ? a.step : a.T - 1 - a.step;
  const int par = a.step & 1;
  const float* hprev = a.hstate + ((size_t)(par * 2 + dir) * B) * H;
  float* hnext = a.hstate + ((size_t)((par ^ 1) * 2 + dir) * B) * H;
  const float* Wh = a.kernel[dir] + (size_t)a.D_in * 4 * H;     // rows D_in.. are the h rows
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.zrec) {                      // recurrent product done by a split-K GEMM launch (H too wide for this loop)
    if (u < H && b < B) {
      const float* zr = a.zrec + ((size_t)dir * B + b) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[g] = zr[g * H + u];
    }
  } else
  for (int k0 = 0; k0 < H; k0 += RK) {
    // Ws[k][g*RU + uu] = Wh[k0+k][g*H + u0+uu]   (RK x 64 = 2048 elems, 8 per thread)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = threadIdx.x + r * 256;
      const int col = e & 63, k = e >> 6;
      const int g = col >> 4, uu = col & 15;
      float v = 0.f;
      if (k0 + k < H && u0 + uu < H) v = Wh[(size_t)(k0 + k) * 4 * H + g * H + u0 + uu];
      Ws[k][col] = v;
    }
    // Hs[bb][k] = hprev[b0+bb][k0+k]   (16 x 32 = 512 elems, 2 per thread)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = threadIdx.x + r * 256;
      const int k = e & 31, bb = e >> 5;
      float v = 0.f;
      if (b0 + bb < B && k0 + k < H) v = hprev[(size_t)(b0 + bb) * H + k0 + k];
      Hs[bb][k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RK; ++k) {
      const float hv = Hs[tb][k];
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[g] = fmaf(hv, Ws[k][g * RU + tu], acc[g]);
    }
    __syncthreads();
  }
  if (u >= H || b >= B) return;
  const size_t cidx = ((size_t)dir * B + b) * H + u;
  const bool active = t < a.seq_len[b];
  const float c_prev = a.cstate[cidx];
  const float h_prev = a.mout ? 0.f : hprev[(size_t)b * H + u];
  const size_t row = (size_t)t * B + b;
  float h_out = 0.f, c_new = c_prev, h_state = h_prev;
  float gi = 0.f, gg = 0.f, gf = 0.f, go = 0.f;
  if (active) {
    const float* Gr = a.G + row * 8 * H + (size_t)dir * 4 * H;
    float zi = Gr[u] + acc[0], zg = Gr[H + u] + acc[1];
    float zf = Gr[2 * H + u] + acc[2] + a.forget_bias, zo = Gr[3 * H + u] + acc[3];
    if (a.use_peephole) { zi += a.wi[dir][u] * c_prev; zf += a.wf[dir][u] * c_prev; }
    gi = sigmoidf_(zi); gg = tanhf_(zg); gf = sigmoidf_(zf);
    c_new = gf * c_prev + gi * gg;
    if (a.cell_clip > 0.f) c_new = fminf(fmaxf(c_new, -a.cell_clip), a.cell_clip);
    if (a.use_peephole) zo += a.wo[dir][u] * c_new;
    go = sigmoidf_(zo);
    h_out = go * tanhf_(c_new);
    h_state = h_out;
  }
  a.cstate[cidx] = c_new;
  if (a.mout) {
    a.mout[cidx] = h_out;
  } else {
    hnext[(size_t)b * H + u] = h_state;
    const size_t oidx = row * 2 * H + (size_t)dir * H + u;
    float yv = h_out;
    if (a.keep_prob < 1.f && active)
      yv = dropout_keep(a.seed, oidx, a.keep_prob) ? h_out / a.keep_prob : 0.f;
    a.y[oidx] = yv;
  }
  if (a.gates) {
    *(float4*)(a.gates + ((row * 2 + dir) * H + u) * 4) = make_float4(gi, gg, gf, go);
    a.cs[(row * 2 + dir) * H + u] = c_new;
    a.hs[(row * 2 + dir) * H + u] = h_out;
  }
}

struct BwdStepArgs {
  int T, B, D_in, H, step;          // step i: fw works on t=T-1-i, bw on t=i
  int use_peephole; float cell_clip, keep_prob; unsigned long long seed;
  const float* kernel[2]; const float* wi[2]; const float* wf[2]; const float* wo[2];
  const int* seq_len;
  const float* dy;                  // [T,B,2H]
  const float* gates; const float* cs;
  float* dG;                        // [T,B,8H]
  float* dcstate;                   // [2][B][H]
  const float* dfinal;              // [4,B,H] d(c_fw,h_fw,c_bw,h_bw) or null
  const float* dhrec;               // [2,B,H] dz_next . Wh^T computed by the caller (wide-H path) or null
  int proj;                         // projection mode: dhrec is the TOTAL dh of this frame (dy and the recurrent
                                    // part went through the projection in the caller)
};

__global__ void __launch_bounds__(256)
lstm_bwd_step_kernel(const BwdStepArgs a) {
  __shared__ float Ws[RU][RK + 1];
  __shared__ float Zs[RB][RK + 1];
  const int dir = blockIdx.z;
  const int u0 = blockIdx.x * RU, b0 = blockIdx.y * RB;
  const int tu = threadIdx.x & 15, tb = threadIdx.x >> 4;
  const int u = u0 + tu, b = b0 + tb;
  const int H = a.H, B = a.B, T = a.T;
  const int t = dir == 0 ? T - 1 - a.step : a.step;
  const int tn = dir == 0 ? t + 1 : t - 1;          // the step processed just before (in BPTT order)
  const float* Wh = a.kernel[dir] + (size_t)a.D_in * 4 * H;
  float acc = 0.f;
  if (a.dhrec) {
    if (u < H && b < B) acc = a.dhrec[((size_t)dir * B + b) * H + u];
  } else if (tn >= 0 && tn < T) {
    const float* dzn = a.dG + (size_t)tn * B * 8 * H + (size_t)dir * 4 * H;   // row b: + b*8H
    for (int k0 = 0; k0 < 4 * H; k0 += RK) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int e = threadIdx.x + r * 256;
        const int k = e & 31, rr = e >> 5;
        float wv = 0.f, zv = 0.f;
        if (u0 + rr < H && k0 + k < 4 * H) wv = Wh[(size_t)(u0 + rr) * 4 * H + k0 + k];
        if (b0 + rr < B && k0 + k < 4 * H) zv = dzn[(size_t)(b0 + rr) * 8 * H + k0 + k];
        Ws[rr][k] = wv;
        Zs[rr][k] = zv;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < RK; ++k) acc = fmaf(Zs[tb][k], Ws[tu][k], acc);
      __syncthreads();
    }
  }
  if (u >= H || b >= B) return;
  const int len = a.seq_len[b];
  const bool active = t < len;
  const size_t row = (size_t)t * B + b;
  float* dz = a.dG + row * 8 * H + (size_t)dir * 4 * H;
  if (!active) {
    dz[u] = 0.f; dz[H + u] = 0.f; dz[2 * H + u] = 0.f; dz[3 * H + u] = 0.f;
    return;
  }
  const bool nb_active = (tn >= 0 && tn < T && tn < len);
  const size_t cidx = ((size_t)dir * B + b) * H + u;
  float dh_in, dc_in;
  if (nb_active) { dh_in = acc; dc_in = a.dcstate[cidx]; }
  else {
    dh_in = a.dfinal ? a.dfinal[((size_t)(dir * 2 + 1) * B + b) * H + u] : 0.f;
    dc_in = a.dfinal ? a.dfinal[((size_t)(dir * 2 + 0) * B + b) * H + u] : 0.f;
  }
  float dh;
  if (a.proj) {
    dh = acc;
  } else {
    const size_t oidx = row * 2 * H + (size_t)dir * H + u;
    float dyv = a.dy[oidx];
    if (a.keep_prob < 1.f) dyv = dropout_keep(a.seed, oidx, a.keep_prob) ? dyv / a.keep_prob : 0.f;
    dh = dyv + dh_in;
  }
  const float4 g4 = *(const float4*)(a.gates + ((row * 2 + dir) * H + u) * 4);
  const float gi = g4.x, gg = g4.y, gf = g4.z, go = g4.w;
  const float c = a.cs[(row * 2 + dir) * H + u];
  const int tp = dir == 0 ? t - 1 : t + 1;           // previous step in forward order
  float c_prev = 0.f;
  if (tp >= 0 && tp < T && tp < len) c_prev = a.cs[(((size_t)tp * B + b) * 2 + dir) * H + u];
  const float tc = tanhf_(c);
  const float dzo = dh * tc * go * (1.f - go);
  float dc = dc_in + dh * go * (1.f - tc * tc);
  if (a.use_peephole) dc += dzo * a.wo[dir][u];
  if (a.cell_clip > 0.f && fabsf(c) >= a.cell_clip) dc = 0.f;   // clip_by_value passes no grad outside
  const float dzi = dc * gg * gi * (1.f - gi);
  const float dzg = dc * gi * (1.f - gg * gg);
  const float dzf = dc * c_prev * gf * (1.f - gf);
  float dc_prev = dc * gf;
  if (a.use_peephole) dc_prev += dzi * a.wi[dir][u] + dzf * a.wf[dir][u];
  a.dcstate[cidx] = dc_prev;
  dz[u] = dzi; dz[H + u] = dzg; dz[2 * H + u] = dzf; dz[3 * H + u] = dzo;
}

// ---- LSTMCell projection (num_proj): per-frame epilogue of the forward pass and prologue of BPTT
// hp_new [2,B,P] = m . Wp (this frame).  Active rows take it, the others keep their state; y gets the
// (dropout-scaled) emitted value, hps the emitted value before dropout.  grid over 2*B*P elements.
__global__ void __launch_bounds__(256)
proj_finalize_kernel(const float* __restrict__ hp_new, const float* __restrict__ hp_prev,
                     float* __restrict__ hp_next, const int* __restrict__ seq_len, int T, int B, int P, int step,
                     float keep_prob, unsigned long long seed, float* __restrict__ y, float* __restrict__ hps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * B * P) return;
  const int p = idx % P, b = (idx / P) % B, dir = idx / (P * B);
  const int t = dir == 0 ? step : T - 1 - step;
  const bool active = t < seq_len[b];
  const float v = hp_new[idx];
  hp_next[idx] = active ? v : hp_prev[idx];
  const size_t oidx = ((size_t)t * B + b) * 2 * P + (size_t)dir * P + p;
  float yv = active ? v : 0.f;
  if (keep_prob < 1.f && active) yv = dropout_keep(seed, oidx, keep_prob) ? v / keep_prob : 0.f;
  y[oidx] = yv;
  if (hps) hps[oidx] = active ? v : 0.f;
}

// d(projected h) of this BPTT frame = dropout'(dy) + recurrent part (or the final-state gradient at the
// first active frame); written for the projection GEMM of this frame and kept for d(projection).
__global__ void __launch_bounds__(256)
proj_bwd_combine_kernel(const float* __restrict__ dy, const float* __restrict__ dhp_rec,
                        const int* __restrict__ seq_len, int T, int B, int P, int step, float keep_prob,
                        unsigned long long seed, float* __restrict__ dhp, float* __restrict__ dhp_all) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * B * P) return;
  const int p = idx % P, b = (idx / P) % B, dir = idx / (P * B);
  const int t = dir == 0 ? T - 1 - step : step;
  const int tn = dir == 0 ? t + 1 : t - 1;
  const int len = seq_len[b];
  const size_t oidx = ((size_t)t * B + b) * 2 * P + (size_t)dir * P + p;
  float v = 0.f;
  if (t < len) {
    float dyv = dy[oidx];
    if (keep_prob < 1.f) dyv = dropout_keep(seed, oidx, keep_prob) ? dyv / keep_prob : 0.f;
    const bool nb_active = step > 0 && tn >= 0 && tn < T && tn < len;
    v = dyv + (nb_active ? dhp_rec[idx] : 0.f);
  }
  dhp[idx] = v;
  dhp_all[oidx] = v;
}

// two products of one shape (forward / backward direction): skinny pair when the batch allows
static int pair_gemm(int transb, int M, int N, int K, const float* A0, const float* A1, int lda, const float* B0,
                     const float* B1, int ldb, float* C0, float* C1, int ldc, cudaStream_t stream) {
  if (M <= 64 && C1 == C0 + (size_t)M * ldc)
    return gemm_skinny_pair(transb, M, N, K, A0, A1, lda, B0, B1, ldb, C0, C1, ldc, stream);
  int rc = gemm_simt(0, transb, M, N, K, 1.f, A0, lda, B0, ldb, 0.f, C0, ldc, nullptr, stream);
  if (rc) return rc;
  return gemm_simt(0, transb, M, N, K, 1.f, A1, lda, B1, ldb, 0.f, C1, ldc, nullptr, stream);
}

// peephole gradients: dwi[u] += sum_{t,b} dz_i * c_prev ; dwf likewise ; dwo += dz_o * c
// grid = (ceil(H/32), slabs, 2 dirs); block 256 = 8 row-lanes x 32 units
__global__ void __launch_bounds__(256)
peephole_grad_kernel(const float* __restrict__ dG, const float* __restrict__ cs,
                     const int* __restrict__ seq_len, int T, int B, int H,
                     float* dwi0, float* dwf0, float* dwo0, float* dwi1, float* dwf1, float* dwo1) {
  __shared__ float sh[3][8][33];
  const int dir = blockIdx.z;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int u = blockIdx.x * 32 + lane;
  float si = 0.f, sf = 0.f, so = 0.f;
  if (u < H) {
    const int64_t rows = (int64_t)T * B;
    for (int64_t r = (int64_t)blockIdx.y * 8 + wy; r < rows; r += (int64_t)gridDim.y * 8) {
      const int t = (int)(r / B), b = (int)(r % B);
      const int len = seq_len[b];
      if (t >= len) continue;
      const float* dz = dG + r * 8 * H + (size_t)dir * 4 * H;
      const float c = cs[(r * 2 + dir) * H + u];
      const int tp = dir == 0 ? t - 1 : t + 1;
      float cp = 0.f;
      if (tp >= 0 && tp < T && tp < len) cp = cs[(((int64_t)tp * B + b) * 2 + dir) * H + u];
      si = fmaf(dz[u], cp, si);
      sf = fmaf(dz[2 * H + u], cp, sf);
      so = fmaf(dz[3 * H + u], c, so);
    }
  }
  sh[0][wy][lane] = si; sh[1][wy][lane] = sf; sh[2][wy][lane] = so;
  __syncthreads();
  if (wy == 0 && u < H) {
#pragma unroll
    for (int j = 1; j < 8; ++j) { si += sh[0][j][lane]; sf += sh[1][j][lane]; so += sh[2][j][lane]; }
    atomicAdd((dir ? dwi1 : dwi0) + u, si);
    atomicAdd((dir ? dwf1 : dwf0) + u, sf);
    atomicAdd((dir ? dwo1 : dwo0) + u, so);
  }
}

// ---------------------------------------------------------------------------
// LSTMCell with num_proj (models/encoders/core/blstm.py:187-255, cell equations
// models/recurrent/layers/lstm.py:166-176): h_t = (o * tanh(c_t)) . W_proj, and it is this projected
// h that recurs and is emitted.  fp32 CUDA-core path: per frame one skinny GEMM pair for the recurrent
// product, the gate-math kernel, one skinny GEMM pair for the projection, one finalize kernel.
// ---------------------------------------------------------------------------
static int proj_forward(const b2_lstm_desc* d, const float* x, const int32_t* seq_len, const b2_lstm_params* fw,
                        const b2_lstm_params* bw, float* y, float* final_state, void* reserve, const Work& w,
                        cudaStream_t stream) {
  const int T = d->T, B = d->B, D = d->D_in, H = d->H, P = d->num_proj, TB = T * B;
  const b2_lstm_params* Pm[2] = {fw, bw};
  B2_CHECK_ARG(fw->projection && bw->projection, "blstm_forward: num_proj without projection weights");
  Reserve r = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (d->need_backward) reserve_layout(d, reserve, &r);
  int rc;
  for (int dir = 0; dir < 2; ++dir) {
    rc = gemm_simt(0, 0, TB, 4 * H, D, 1.f, x, D, Pm[dir]->kernel, 4 * H, 0.f, w.G + (size_t)dir * 4 * H, 8 * H,
                   Pm[dir]->bias, stream);
    if (rc) return rc;
  }
  B2_CUDA(cudaMemsetAsync(w.hstate, 0, (size_t)2 * 2 * B * P * sizeof(float), stream));
  B2_CUDA(cudaMemsetAsync(w.cstate, 0, (size_t)2 * B * H * sizeof(float), stream));
  StepArgs a;
  a.T = T; a.B = B; a.D_in = D; a.H = H;
  a.use_peephole = d->use_peephole; a.forget_bias = d->forget_bias; a.cell_clip = d->cell_clip;
  a.keep_prob = 1.f; a.seed = d->dropout_seed;            // dropout is applied after the projection
  for (int dir = 0; dir < 2; ++dir) {
    a.kernel[dir] = Pm[dir]->kernel; a.wi[dir] = Pm[dir]->w_i_diag; a.wf[dir] = Pm[dir]->w_f_diag;
    a.wo[dir] = Pm[dir]->w_o_diag;
  }
  a.seq_len = seq_len; a.G = w.G; a.hstate = w.hstate; a.cstate = w.cstate; a.y = y;
  a.gates = r.gates; a.cs = r.cs; a.hs = r.hs; a.zrec = w.zrec; a.mout = w.mcur;
  dim3 grid(cdiv(H, RU), cdiv(B, RB), 2);
  for (int i = 0; i < T; ++i) {
    a.step = i;
    float* hp_prev = w.hstate + (size_t)((i & 1) * 2) * B * P;          // [dir][B][P]
    float* hp_next = w.hstate + (size_t)(((i & 1) ^ 1) * 2) * B * P;
    rc = pair_gemm(0, B, 4 * H, P, hp_prev, hp_prev + (size_t)B * P, P, fw->kernel + (size_t)D * 4 * H,
                   bw->kernel + (size_t)D * 4 * H, 4 * H, w.zrec, w.zrec + (size_t)B * 4 * H, 4 * H, stream);
    if (rc) return rc;
    lstm_fwd_step_kernel<<<grid, 256, 0, stream>>>(a);
    rc = pair_gemm(0, B, P, H, w.mcur, w.mcur + (size_t)B * H, H, fw->projection, bw->projection, P, w.hpnew,
                   w.hpnew + (size_t)B * P, P, stream);
    if (rc) return rc;
    proj_finalize_kernel<<<cdiv((int64_t)2 * B * P, 256), 256, 0, stream>>>(w.hpnew, hp_prev, hp_next, seq_len, T, B, P, i,
                                                                           d->keep_prob, d->dropout_seed, y, r.hps);
  }
  count_launches(2 * T - 1);
  B2_LAUNCH_CHECK();
  if (final_state) {      // c_fw [B,H], h_fw [B,P], c_bw [B,H], h_bw [B,P]
    const float* hfin = w.hstate + (size_t)((T & 1) * 2) * B * P;
    float* o = final_state;
    B2_CUDA(cudaMemcpyAsync(o, w.cstate, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, stream)); o += (size_t)B * H;
    B2_CUDA(cudaMemcpyAsync(o, hfin, (size_t)B * P * 4, cudaMemcpyDeviceToDevice, stream)); o += (size_t)B * P;
    B2_CUDA(cudaMemcpyAsync(o, w.cstate + (size_t)B * H, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, stream)); o += (size_t)B * H;
    B2_CUDA(cudaMemcpyAsync(o, hfin + (size_t)B * P, (size_t)B * P * 4, cudaMemcpyDeviceToDevice, stream));
  }
  return B2_OK;
}

static int proj_backward(const b2_lstm_desc* d, const float* x, const int32_t* seq_len, const b2_lstm_params* fw,
                         const b2_lstm_params* bw, const float* dy, const void* reserve, float* dx,
                         const b2_lstm_grads* g_fw, const b2_lstm_grads* g_bw, const Work& w, cudaStream_t stream) {
  const int T = d->T, B = d->B, D = d->D_in, H = d->H, P = d->num_proj, TB = T * B;
  const b2_lstm_params* Pm[2] = {fw, bw};
  const b2_lstm_grads* Gr[2] = {g_fw, g_bw};
  B2_CHECK_ARG(fw->projection && bw->projection && g_fw->projection && g_bw->projection,
               "blstm_backward: num_proj without projection weights / gradients");
  Reserve r;
  reserve_layout(d, (void*)reserve, &r);
  int rc;
  B2_CUDA(cudaMemsetAsync(w.cstate, 0, (size_t)2 * B * H * sizeof(float), stream));
  BwdStepArgs a;
  a.T = T; a.B = B; a.D_in = D; a.H = H;
  a.use_peephole = d->use_peephole; a.cell_clip = d->cell_clip; a.keep_prob = 1.f; a.seed = d->dropout_seed;
  for (int dir = 0; dir < 2; ++dir) {
    a.kernel[dir] = Pm[dir]->kernel; a.wi[dir] = Pm[dir]->w_i_diag; a.wf[dir] = Pm[dir]->w_f_diag;
    a.wo[dir] = Pm[dir]->w_o_diag;
  }
  a.seq_len = seq_len; a.dy = dy; a.gates = r.gates; a.cs = r.cs; a.dG = w.G; a.dcstate = w.cstate;
  a.dfinal = nullptr; a.dhrec = w.mcur; a.proj = 1;
  float* dhp_rec = w.zrec;                 // [2,B,P]
  float* dhp = w.hpnew;                    // [2,B,P]
  dim3 grid(cdiv(H, RU), cdiv(B, RB), 2);
  for (int i = 0; i < T; ++i) {
    a.step = i;
    if (i > 0) {
      const float* dz_fw = w.G + (size_t)(T - i) * B * 8 * H;
      const float* dz_bw = w.G + (size_t)(i - 1) * B * 8 * H + (size_t)4 * H;
      rc = pair_gemm(1, B, P, 4 * H, dz_fw, dz_bw, 8 * H, fw->kernel + (size_t)D * 4 * H,
                     bw->kernel + (size_t)D * 4 * H, 4 * H, dhp_rec, dhp_rec + (size_t)B * P, P, stream);
      if (rc) return rc;
    }
    proj_bwd_combine_kernel<<<cdiv((int64_t)2 * B * P, 256), 256, 0, stream>>>(dy, dhp_rec, seq_len, T, B, P, i, d->keep_prob,
                                                                              d->dropout_seed, dhp, w.dhp_all);
    rc = pair_gemm(1, B, H, P, dhp, dhp + (size_t)B * P, P, fw->projection, bw->projection, P, w.mcur,
                   w.mcur + (size_t)B * H, H, stream);
    if (rc) return rc;
    lstm_bwd_step_kernel<<<grid, 256, 0, stream>>>(a);
  }
  count_launches(2 * T - 1);
  B2_LAUNCH_CHECK();
  for (int dir = 0; dir < 2; ++dir) {
    rc = b2_colsum(w.G + (size_t)dir * 4 * H, TB, 4 * H, 8 * H, Gr[dir]->bias, 1, (b2_stream_t)stream);
    if (rc) return rc;
  }
  if (d->use_peephole) {
    int slabs = cdiv(TB, 64); if (slabs > 128) slabs = 128;
    dim3 pg(cdiv(H, 32), slabs, 2);
    peephole_grad_kernel<<<pg, 256, 0, stream>>>(w.G, r.cs, seq_len, T, B, H, g_fw->w_i_diag, g_fw->w_f_diag,
                                                g_fw->w_o_diag, g_bw->w_i_diag, g_bw->w_f_diag, g_bw->w_o_diag);
    B2_LAUNCH_CHECK();
  }
  for (int dir = 0; dir < 2; ++dir) {
    const float* dGd = w.G + (size_t)dir * 4 * H;
    if (dx) {
      rc = gemm_simt(0, 1, TB, D, 4 * H, 1.f, dGd, 8 * H, Pm[dir]->kernel, 4 * H, dir == 0 ? 0.f : 1.f, dx, D, nullptr, stream);
      if (rc) return rc;
    }
    rc = gemm_simt(1, 0, D, 4 * H, TB, 1.f, x, D, dGd, 8 * H, 1.f, Gr[dir]->kernel, 4 * H, nullptr, stream);
    if (rc) return rc;
    if (T > 1) {        // d(Wh) [P,4H] += Hp_prev^T . dG  (projected h of the neighbouring frame)
      const float* hp_a = r.hps + (dir == 0 ? 0 : (size_t)B * 2 * P) + (size_t)dir * P;
      const float* dG_h = dGd + (dir == 0 ? (size_t)B * 8 * H : 0);
      rc = gemm_simt(1, 0, P, 4 * H, (T - 1) * B, 1.f, hp_a, 2 * P, dG_h, 8 * H, 1.f,
                     Gr[dir]->kernel + (size_t)D * 4 * H, 4 * H, nullptr, stream);
      if (rc) return rc;
    }
    // d(W_proj) [H,P] += M^T . d(projected h)
    rc = gemm_simt(1, 0, H, P, TB, 1.f, r.hs + (size_t)dir * H, 2 * H, w.dhp_all + (size_t)dir * P, 2 * P, 1.f,
                   Gr[dir]->projection, P, nullptr, stream);
    if (rc) return rc;
  }
  return B2_OK;
}

}  // namespace b2

using namespace b2;

static int check_desc(const b2_lstm_desc* d) {
  B2_CHECK_ARG(d != nullptr, "lstm: null descriptor");
  B2_CHECK_ARG(d->T > 0 && d->B > 0 && d->D_in > 0 && d->H > 0, "lstm: bad shape T=%d B=%d D=%d H=%d",
               d->T, d->B, d->D_in, d->H);
  B2_CHECK_ARG(d->keep_prob > 0.f && d->keep_prob <= 1.f, "lstm: keep_prob %f out of (0,1]",
               d->keep_prob);
  B2_CHECK_ARG(d->precision == B2_PREC_FP32 || d->precision == B2_PREC_BF16,
               "lstm: unknown precision %d", d->precision);
  B2_CHECK_ARG(d->num_proj >= 0 && d->num_proj <= 4 * d->H, "lstm: num_proj %d out of range", d->num_proj);
  return B2_OK;
}

extern "C" size_t b2_blstm_reserve_bytes(const b2_lstm_desc* d) {
  return d ? reserve_layout(d, nullptr, nullptr) : 0;
}
extern "C" size_t b2_blstm_workspace_bytes(const b2_lstm_desc* d) {
  if (!d) return 0;
  if (tc_layer_supported(d)) return tc_layer_workspace_bytes(d);
  return work_layout(d, nullptr, nullptr);
}

extern "C" void b2_blstm_profile_enable(int on) { tc_profile_enable(on); }
extern "C" int b2_blstm_profile_last_ms(float* fwd_ms, float* bwd_ms) {
  return tc_profile_last_ms(fwd_ms, bwd_ms);
}

extern "C" int b2_blstm_backward_join(b2_stream_t stream_) {
  return tc_backward_join((cudaStream_t)stream_);
}

extern "C" int b2_blstm_backward_side_wait(b2_stream_t stream_) {
  return tc_backward_side_wait((cudaStream_t)stream_);
}

extern "C" int b2_blstm_layer_path(const b2_lstm_desc* d) {
  if (!d) return 0;
  if (tc_layer_supported(d)) return 1;
  if (wide_rec_supported(d)) return 2;
  return 0;
}

extern "C" const void* b2_blstm_reserve_y_lp(const b2_lstm_desc* d, const void* reserve) {
  if (!d || !reserve || !(tc_layer_supported(d) || wide_rec_supported(d))) return nullptr;
  Reserve r;
  reserve_layout(d, (void*)reserve, &r);
  return r.y_lp;
}

extern "C" int b2_blstm_layer_forward(const b2_lstm_desc* d, const float* x, const void* x_lp,
                                      const int32_t* seq_len, const b2_lstm_params* fw,
                                      const b2_lstm_params* bw, float* y, float* final_state,
                                      void* reserve, void* workspace, size_t workspace_bytes,
                                      b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_desc(d);
  if (rc) return rc;
  if (tc_layer_supported(d)) {
    B2_CHECK_ARG(x && seq_len && fw && bw && y && workspace, "blstm_forward: null pointer");
    B2_CHECK_ARG(reserve, "blstm_forward(bf16): reserve required");
    return tc_layer_forward(d, x, (const __nv_bfloat16*)x_lp, seq_len, fw, bw, y, final_state,
                            reserve, workspace, workspace_bytes, stream);
  }
  B2_CHECK_ARG(x && seq_len && fw && bw && y && workspace, "blstm_forward: null pointer");
  B2_CHECK_ARG(!d->need_backward || reserve, "blstm_forward: need_backward without reserve");
  B2_CHECK_ARG(!d->use_peephole || (fw->w_i_diag && fw->w_f_diag && fw->w_o_diag && bw->w_i_diag &&
                                    bw->w_f_diag && bw->w_o_diag),
               "blstm_forward: peephole weights missing");
  Work w;
  const size_t need = work_layout(d, workspace, &w);
  if (workspace_bytes < need) { set_error("blstm_forward: workspace %zu < %zu", workspace_bytes, need); return B2_ERR_WORKSPACE; }
  if (d->num_proj > 0) return proj_forward(d, x, seq_len, fw, bw, y, final_state, reserve, w, stream);
  Reserve r = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (d->need_backward) reserve_layout(d, reserve, &r);
  const int T = d->T, B = d->B, D = d->D_in, H = d->H;
  const int TB = T * B;
  const b2_lstm_params* P[2] = {fw, bw};

  // 1. input projection, both directions into G [TB, 8H]
  for (int dir = 0; dir < 2; ++dir) {
    float* Gd = w.G + (size_t)dir * 4 * H;
    if (d->precision == B2_PREC_BF16) {
      // the caller's bf16 shadow of x (the layer below's emitted output) saves the cast pass
      const bool have_lp = x_lp != nullptr && (D % 8) == 0;
      const int ldx = have_lp ? D : (int)pad8z(D);
      const __nv_bfloat16* xa = have_lp ? (const __nv_bfloat16*)x_lp : w.xb;
      if (dir == 0 && !have_lp) { rc = cast_f32_bf16(x, TB, D, D, w.xb, ldx, stream); if (rc) return rc; }
      __nv_bfloat16* wb = w.wb + (size_t)dir * (D + 0) * 4 * H;
      rc = cast_f32_bf16(P[dir]->kernel, D, 4 * H, 4 * H, wb, 4 * H, stream);
      if (rc) return rc;
      rc = gemm_bf16_tc(0, 1, TB, 4 * H, D, 1.f, xa, ldx, wb, 4 * H, Gd, 8 * H, P[dir]->bias,
                        0 /*store f32*/, 0, stream);
    } else {
      rc = gemm_simt(0, 0, TB, 4 * H, D, 1.f, x, D, P[dir]->kernel, 4 * H, 0.f, Gd, 8 * H,
                     P[dir]->bias, stream);
    }
    if (rc) return rc;
  }
  // 2. recurrence
  if (wide_rec_supported(d))        // wide layer (config 4: H = 1024): ONE cooperative launch, weights register-resident
    return wide_rec_forward(d, fw, bw, seq_len, w.G, y, r.gates, r.cs, r.hs_lp, r.y_lp, final_state, w.wide, stream);
  B2_CUDA(cudaMemsetAsync(w.hstate, 0, (size_t)2 * 2 * B * H * sizeof(float), stream));
  B2_CUDA(cudaMemsetAsync(w.cstate, 0, (size_t)2 * B * H * sizeof(float), stream));
  StepArgs a;
  a.T = T; a.B = B; a.D_in = D; a.H = H;
  a.use_peephole = d->use_peephole; a.forget_bias = d->forget_bias; a.cell_clip = d->cell_clip;
  a.keep_prob = d->keep_prob; a.seed = d->dropout_seed;
  for (int dir = 0; dir < 2; ++dir) {
    a.kernel[dir] = P[dir]->kernel; a.wi[dir] = P[dir]->w_i_diag; a.wf[dir] = P[dir]->w_f_diag;
    a.wo[dir] = P[dir]->w_o_diag;
  }
  a.seq_len = seq_len; a.G = w.G; a.hstate = w.hstate; a.cstate = w.cstate; a.y = y;
  a.gates = r.gates; a.cs = r.cs; a.hs = r.hs; a.mout = nullptr;
  // Wide layers (H > 512: the cluster-resident tcgen05 recurrence does not hold them and the serial-K
  // loop of the step kernel takes 80 us per frame at H=1024): the recurrent product h_prev . Wh becomes
  // one split-K skinny GEMM per direction and frame, the step kernel only does the gate math.
  const bool wide = H > 512 && B <= 64;
  a.zrec = wide ? w.zrec : nullptr;
  dim3 grid(cdiv(H, RU), cdiv(B, RB), 2);
  for (int i = 0; i < T; ++i) {
    a.step = i;
    if (wide) {
      const float* hprev = w.hstate + ((size_t)((i & 1) * 2) * B) * H;          // [dir][B][H], this parity
      rc = gemm_skinny_pair(0, B, 4 * H, H, hprev, hprev + (size_t)B * H, H, P[0]->kernel + (size_t)D * 4 * H,
                            P[1]->kernel + (size_t)D * 4 * H, 4 * H, w.zrec, w.zrec + (size_t)B * 4 * H, 4 * H, stream);
      if (rc) return rc;
    }
    lstm_fwd_step_kernel<<<grid, 256, 0, stream>>>(a);
  }
  count_launches(T - 1);
  B2_LAUNCH_CHECK();
  if (final_state) {
    // hstate parity after T steps is T&1; layout [c_fw, h_fw, c_bw, h_bw] each [B,H]
    const float* hfin = w.hstate + (size_t)((T & 1) * 2) * B * H;
    const size_t n = (size_t)B * H * sizeof(float);
    B2_CUDA(cudaMemcpyAsync(final_state, w.cstate, n, cudaMemcpyDeviceToDevice, stream));
    B2_CUDA(cudaMemcpyAsync(final_state + (size_t)B * H, hfin, n, cudaMemcpyDeviceToDevice, stream));
    B2_CUDA(cudaMemcpyAsync(final_state + (size_t)2 * B * H, w.cstate + (size_t)B * H, n, cudaMemcpyDeviceToDevice, stream));
    B2_CUDA(cudaMemcpyAsync(final_state + (size_t)3 * B * H, hfin + (size_t)B * H, n, cudaMemcpyDeviceToDevice, stream));
  }
  return B2_OK;
}

extern "C" int b2_blstm_layer_backward(const b2_lstm_desc* d, const float* x, const void* x_lp,
                                       const int32_t* seq_len, const b2_lstm_params* fw,
                                       const b2_lstm_params* bw, const float* dy,
                                       const void* reserve, float* dx, const b2_lstm_grads* g_fw,
                                       const b2_lstm_grads* g_bw, void* workspace,
                                       size_t workspace_bytes, b2_stream_t stream_) {
  return b2_blstm_layer_backward_ex(d, x, x_lp, seq_len, fw, bw, dy, nullptr, reserve, dx, g_fw, g_bw,
                                    workspace, workspace_bytes, stream_);
}

extern "C" int b2_blstm_layer_backward_ex(const b2_lstm_desc* d, const float* x, const void* x_lp,
                                          const int32_t* seq_len, const b2_lstm_params* fw,
                                          const b2_lstm_params* bw, const float* dy,
                                          const float* d_final_state, const void* reserve,
                                          float* dx, const b2_lstm_grads* g_fw,
                                          const b2_lstm_grads* g_bw, void* workspace,
                                          size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_desc(d);
  if (rc) return rc;
  if (tc_layer_supported(d)) {
    B2_CHECK_ARG(x && seq_len && fw && bw && dy && reserve && g_fw && g_bw && workspace,
                 "blstm_backward: null pointer");
    return tc_layer_backward(d, x, (const __nv_bfloat16*)x_lp, seq_len, fw, bw, dy, d_final_state, reserve, dx,
                             g_fw, g_bw, workspace, workspace_bytes, stream);
  }
  B2_CHECK_ARG(x && seq_len && fw && bw && dy && reserve && g_fw && g_bw && workspace,
               "blstm_backward: null pointer");
  if (d->dy_premasked || (d->dx_keep_prob > 0.f && d->dx_keep_prob < 1.f)) {
    set_error("blstm_backward: dx_keep_prob / dy_premasked need the tcgen05 layer path (b2_blstm_layer_path() == 1)");
    return B2_ERR_UNSUPPORTED;
  }
  Work w;
  const size_t need = work_layout(d, workspace, &w);
  if (workspace_bytes < need) { set_error("blstm_backward: workspace %zu < %zu", workspace_bytes, need); return B2_ERR_WORKSPACE; }
  if (d->num_proj > 0) {
    B2_CHECK_ARG(!d_final_state, "blstm_backward: d_final_state with num_proj is not built");
    return proj_backward(d, x, seq_len, fw, bw, dy, reserve, dx, g_fw, g_bw, w, stream);
  }
  Reserve r;
  reserve_layout(d, (void*)reserve, &r);
  const int T = d->T, B = d->B, D = d->D_in, H = d->H;
  const int TB = T * B;
  const b2_lstm_params* P[2] = {fw, bw};
  const b2_lstm_grads* Gr[2] = {g_fw, g_bw};

  // 1. BPTT recurrence -> dG [TB, 8H]
  B2_CUDA(cudaMemsetAsync(w.cstate, 0, (size_t)2 * B * H * sizeof(float), stream));
  BwdStepArgs a;
  a.T = T; a.B = B; a.D_in = D; a.H = H;
  a.use_peephole = d->use_peephole; a.cell_clip = d->cell_clip; a.keep_prob = d->keep_prob;
  a.seed = d->dropout_seed;
  for (int dir = 0; dir < 2; ++dir) {
    a.kernel[dir] = P[dir]->kernel; a.wi[dir] = P[dir]->w_i_diag; a.wf[dir] = P[dir]->w_f_diag;
    a.wo[dir] = P[dir]->w_o_diag;
  }
  a.seq_len = seq_len; a.dy = dy; a.gates = r.gates; a.cs = r.cs; a.dG = w.G; a.dcstate = w.cstate;
  a.dfinal = d_final_state; a.proj = 0;
  const bool wide = H > 512 && B <= 64;            // see the forward pass
  a.dhrec = wide ? w.zrec : nullptr;
  dim3 grid(cdiv(H, RU), cdiv(B, RB), 2);
  const bool resident = wide_rec_supported(d);     // config 4: ONE cooperative launch for the whole BPTT recurrence
  if (resident) {
    rc = wide_rec_backward(d, fw, bw, seq_len, dy, r.gates, r.cs, d_final_state, w.G, w.gb, w.wide, stream);
    if (rc) return rc;
  }
  for (int i = 0; i < (resident ? 0 : T); ++i) {
    a.step = i;
    if (wide) {
      if (i == 0) B2_CUDA(cudaMemsetAsync(w.zrec, 0, (size_t)2 * B * H * sizeof(float), stream));
      else {
        // the frame processed one BPTT step earlier: fw t+1 = T-i, bw t-1 = i-1
        const float* dz_fw = w.G + (size_t)(T - i) * B * 8 * H;
        const float* dz_bw = w.G + (size_t)(i - 1) * B * 8 * H + (size_t)4 * H;
        rc = gemm_skinny_pair(1, B, H, 4 * H, dz_fw, dz_bw, 8 * H, P[0]->kernel + (size_t)D * 4 * H,
                              P[1]->kernel + (size_t)D * 4 * H, 4 * H, w.zrec, w.zrec + (size_t)B * H, H, stream);
        if (rc) return rc;
      }
    }
    lstm_bwd_step_kernel<<<grid, 256, 0, stream>>>(a);
  }
  if (!resident) { count_launches(T - 1); B2_LAUNCH_CHECK(); }

  // 2. bias + peephole reductions
  for (int dir = 0; dir < 2; ++dir) {
    rc = b2_colsum(w.G + (size_t)dir * 4 * H, TB, 4 * H, 8 * H, Gr[dir]->bias, 1, stream_);
    if (rc) return rc;
  }
  if (d->use_peephole) {
    int slabs = cdiv(TB, 64); if (slabs > 128) slabs = 128;
    dim3 pg(cdiv(H, 32), slabs, 2);
    peephole_grad_kernel<<<pg, 256, 0, stream>>>(w.G, r.cs, seq_len, T, B, H, g_fw->w_i_diag,
                                                g_fw->w_f_diag, g_fw->w_o_diag, g_bw->w_i_diag,
                                                g_bw->w_f_diag, g_bw->w_o_diag);
    B2_LAUNCH_CHECK();
  }

  // 3. time-batched GEMMs
  const bool tc = d->precision == B2_PREC_BF16;
  if (tc && !resident) {          // the grid-resident BPTT kernel has written the bf16 copy itself
    rc = cast_f32_bf16(w.G, TB, 8 * H, 8 * H, w.gb, 8 * H, stream);
    if (rc) return rc;
  }
  const bool have_lp = tc && x_lp != nullptr && (D % 8) == 0;
  for (int dir = 0; dir < 2; ++dir) {
    const float* dGd = w.G + (size_t)dir * 4 * H;
    const __nv_bfloat16* dGb = w.gb + (size_t)dir * 4 * H;
    // dX (+)= dG_dir . Wx_dir^T            [TB,4H] x [4H,D]
    if (dx) {
      if (tc) {
        __nv_bfloat16* wb = w.wb + (size_t)dir * D * 4 * H;
        rc = cast_f32_bf16(P[dir]->kernel, D, 4 * H, 4 * H, wb, 4 * H, stream);
        if (rc) return rc;
        if (dir == 0) B2_CUDA(cudaMemsetAsync(dx, 0, (size_t)TB * D * sizeof(float), stream));
        rc = gemm_bf16_tc(0, 0, TB, D, 4 * H, 1.f, dGb, 8 * H, wb, 4 * H, dx, D, nullptr,
                          1 /*atomic*/, 1, stream);
      } else {
        rc = gemm_simt(0, 1, TB, D, 4 * H, 1.f, dGd, 8 * H, P[dir]->kernel, 4 * H,
                       dir == 0 ? 0.f : 1.f, dx, D, nullptr, stream);
      }
      if (rc) return rc;
    }
    // dWx_dir += X^T . dG_dir              [D,TB] x [TB,4H]
    // dWh_dir += Hprev^T . dG_dir          [H,(T-1)B] x [(T-1)B,4H]   (hs shifted one step)
    const float* hs_a = r.hs + (dir == 0 ? 0 : (size_t)B * 2 * H) + (size_t)dir * H;
    const float* dG_h = dGd + (dir == 0 ? (size_t)B * 8 * H : 0);
    if (tc) {
      const int ldx = have_lp ? D : (int)pad8z(D);
      const __nv_bfloat16* xa = have_lp ? (const __nv_bfloat16*)x_lp : w.xb;
      if (!have_lp) { rc = cast_f32_bf16(x, TB, D, D, w.xb, ldx, stream); if (rc) return rc; }
      rc = gemm_bf16_tc(1, 1, D, 4 * H, TB, 1.f, xa, ldx, dGb, 8 * H, Gr[dir]->kernel, 4 * H,
                        nullptr, 1, 0, stream);
      if (rc) return rc;
      if (T > 1) {
        const __nv_bfloat16* gb = dGb + (dir == 0 ? (size_t)B * 8 * H : 0);
        if (r.hs_lp) {
          // the recurrence kernel kept h (before dropout) in bf16: [TB, 2H], this direction's columns, one step shifted
          const __nv_bfloat16* ha = r.hs_lp + (size_t)dir * H + (dir == 0 ? 0 : (size_t)B * 2 * H);
          rc = gemm_bf16_tc(1, 1, H, 4 * H, (T - 1) * B, 1.f, ha, 2 * H, gb, 8 * H,
                            Gr[dir]->kernel + (size_t)D * 4 * H, 4 * H, nullptr, 1, 0, stream);
        } else {
          // bf16 copy of this direction's hs columns, compact [TB, H]
          rc = cast_f32_bf16(r.hs + (size_t)dir * H, TB, H, 2 * H, w.xb, H, stream);
          if (rc) return rc;
          const __nv_bfloat16* ha = w.xb + (dir == 0 ? 0 : (size_t)B * H);
          rc = gemm_bf16_tc(1, 1, H, 4 * H, (T - 1) * B, 1.f, ha, H, gb, 8 * H,
                            Gr[dir]->kernel + (size_t)D * 4 * H, 4 * H, nullptr, 1, 0, stream);
        }
        if (rc) return rc;
      }
    } else {
      rc = gemm_simt(1, 0, D, 4 * H, TB, 1.f, x, D, dGd, 8 * H, 1.f, Gr[dir]->kernel, 4 * H,
                     nullptr, stream);
      if (rc) return rc;
      if (T > 1) {
        rc = gemm_simt(1, 0, H, 4 * H, (T - 1) * B, 1.f, hs_a, 2 * H, dG_h, 8 * H, 1.f,
                       Gr[dir]->kernel + (size_t)D * 4 * H, 4 * H, nullptr, stream);
        if (rc) return rc;
      }
    }
  }
  return B2_OK;
}
