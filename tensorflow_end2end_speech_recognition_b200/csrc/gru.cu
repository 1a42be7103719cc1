// Bidirectional GRU layer (tf.contrib.rnn.GRUCell under bidirectional_dynamic_rnn / MultiRNNCell + dynamic_rnn),
// forward and BPTT, fp32.  Replaces the cells of models/encoders/core/gru.py:52-73 (GRUEncoder) and :128-160
// (BGRUEncoder); TF 1.x GRUCell arithmetic:
//   [r, u] = sigmoid([x, h] . W_gates + b_gates)          (W_gates [(D+H), 2H], b_gates initialised to 1)
//   c      = tanh([x, r*h] . W_cand + b_cand)             (W_cand  [(D+H), H])
//   h'     = u * h + (1 - u) * c
// with dynamic_rnn's sequence_length semantics (state carried, output zero past the length) and DropoutWrapper on
// the emitted output only.
//
// A widening row, not the headline path: the time-batched products (input projections, weight and
// input gradients) are single GEMMs, the recurrence is a per-frame sequence of launches -- two skinny GEMM pairs
// (both directions per launch: h.U_gates, then (r*h).U_cand -- the candidate needs r first) and two gate kernels per
// frame, mirrored in BPTT.  No persistent kernel: the two dependent products per step would need a grid-wide
// exchange twice per frame.
#include "common.cuh"
#include "lstm_internal.cuh"

namespace b2 {
namespace {

struct GruWork {
  float* gru;     // [TB][2][2H] gate pre-activations from x (+bias); backward: d(pre-activations)
  float* gc;      // [TB][2][H]  candidate pre-activation from x (+bias); backward: its gradient
  float* rhs;     // [TB][2][H]  r * h_prev (backward: operand of the candidate weight gradient)
  float* zru;     // [2][B][2H]  recurrent part of the gate pre-activations / backward: d(gates) of this frame
  float* rh;      // [2][B][H]   r * h_prev of this frame / backward: d(candidate pre-activation)
  float* zc;      // [2][B][H]   recurrent part of the candidate / backward: d(r*h)
  float* hstate;  // [2][B][H]   carried state / backward: carried dh
  float* du;      // [2][B][H]   backward scratch
  float* dhd;     // [2][B][H]   backward scratch (direct path of dh)
  float* dhp;     // [2][B][H]   backward: dz_gates . U_gates^T
};

size_t gru_work_layout(const b2_gru_desc* d, void* base, GruWork* w) {
  const size_t TB = (size_t)d->T * d->B, H = d->H, B = d->B;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n * sizeof(float), 256); return o; };
  const size_t o0 = take(TB * 4 * H), o1 = take(TB * 2 * H), o2 = take(TB * 2 * H);
  const size_t o3 = take(2 * B * 2 * H), o4 = take(2 * B * H), o5 = take(2 * B * H), o6 = take(2 * B * H);
  const size_t o7 = take(2 * B * H), o8 = take(2 * B * H), o9 = take(2 * B * H);
  if (w) {
    char* p = (char*)base;
    w->gru = (float*)(p + o0); w->gc = (float*)(p + o1); w->rhs = (float*)(p + o2); w->zru = (float*)(p + o3);
    w->rh = (float*)(p + o4); w->zc = (float*)(p + o5); w->hstate = (float*)(p + o6); w->du = (float*)(p + o7);
    w->dhd = (float*)(p + o8); w->dhp = (float*)(p + o9);
  }
  return off;
}

struct GruReserve { float* r; float* u; float* c; float* hs; };   // each [T][B][2][H]

size_t gru_reserve_layout(const b2_gru_desc* d, void* base, GruReserve* r) {
  const size_t n = (size_t)d->T * d->B * 2 * d->H;
  const size_t each = align_up(n * sizeof(float), 256);
  if (r) {
    char* p = (char*)base;
    r->r = (float*)p; r->u = (float*)(p + each); r->c = (float*)(p + 2 * each); r->hs = (float*)(p + 3 * each);
  }
  return 4 * each;
}

// frame processed at step s by direction dir
__device__ __forceinline__ int frame_of(int s, int dir, int T) { return dir ? T - 1 - s : s; }

// ---- forward, stage 1: r, u and r*h
__global__ void __launch_bounds__(256)
gru_fwd_gates_kernel(int s, int T, int B, int H, const float* __restrict__ gru, const float* __restrict__ zru,
                     const float* __restrict__ hstate, float* __restrict__ R, float* __restrict__ U,
                     float* __restrict__ rh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * H) return;
  const int u = i % H, b = (i / H) % B, dir = i / (H * B);
  const int td = frame_of(s, dir, T);
  const size_t row = (size_t)td * B + b;
  const float* g = gru + (row * 2 + dir) * 2 * H;
  const float* z = zru + ((size_t)dir * B + b) * 2 * H;
  const float r = sigmoidf_(g[u] + z[u]);
  const float uu = sigmoidf_(g[H + u] + z[H + u]);
  const size_t cell = (row * 2 + dir) * H + u;
  R[cell] = r; U[cell] = uu;
  rh[i] = r * hstate[i];
}

// ---- forward, stage 2: candidate, new state, emitted output
__global__ void __launch_bounds__(256)
gru_fwd_out_kernel(int s, int T, int B, int H, const float* __restrict__ gc, const float* __restrict__ zc,
                   const float* __restrict__ U, const int* __restrict__ seq_len, float keep, unsigned long long seed,
                   float* __restrict__ hstate, float* __restrict__ C, float* __restrict__ HS, float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * H) return;
  const int u = i % H, b = (i / H) % B, dir = i / (H * B);
  const int td = frame_of(s, dir, T);
  const size_t row = (size_t)td * B + b;
  const size_t cell = (row * 2 + dir) * H + u;
  const float c = tanhf_(gc[cell] + zc[i]);
  const float uu = U[cell];
  const float h = hstate[i];
  const bool active = td < seq_len[b];
  const float hn = active ? fmaf(uu, h - c, c) : h;          // u*h + (1-u)*c
  hstate[i] = hn;
  C[cell] = c; HS[cell] = hn;
  float out = active ? hn : 0.f;
  const size_t oidx = row * 2 * H + (size_t)dir * H + u;
  if (keep < 1.f && active) out = dropout_keep(seed, oidx, keep) ? out / keep : 0.f;
  y[oidx] = out;
}

// state before frame td of direction dir: HS of the frame visited one step earlier (0 at the first step)
__device__ __forceinline__ float h_prev_of(const float* HS, int td, int dir, int T, int B, int H, int b, int u) {
  const int tq = dir ? td + 1 : td - 1;
  if (tq < 0 || tq >= T) return 0.f;
  return HS[(((size_t)tq * B + b) * 2 + dir) * H + u];
}

// ---- BPTT, stage 1: through h' = u*h + (1-u)*c and tanh
__global__ void __launch_bounds__(256)
gru_bwd_out_kernel(int s, int T, int B, int H, const float* __restrict__ dy, const float* __restrict__ U,
                   const float* __restrict__ C, const float* __restrict__ HS, const int* __restrict__ seq_len,
                   float keep, unsigned long long seed, int first, const float* __restrict__ dhp,
                   float* __restrict__ dh, float* __restrict__ du, float* __restrict__ dhd,
                   float* __restrict__ dzc_s, float* __restrict__ dzc_all) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * H) return;
  const int u = i % H, b = (i / H) % B, dir = i / (H * B);
  // BPTT visits the frames in the reverse of the forward order
  const int td = frame_of(T - 1 - s, dir, T);
  const size_t row = (size_t)td * B + b;
  const size_t cell = (row * 2 + dir) * H + u;
  // finish the previous BPTT step: dh += dz_gates . U_gates^T (only where that frame was active: dhp is 0 otherwise)
  float dhc = dh[i] + (first ? 0.f : dhp[i]);
  const bool active = td < seq_len[b];
  float dzc = 0.f, duv = 0.f, dd = dhc;
  if (active) {
    const size_t oidx = row * 2 * H + (size_t)dir * H + u;
    float g = dy[oidx];
    if (keep < 1.f) g = dropout_keep(seed, oidx, keep) ? g / keep : 0.f;
    const float dht = g + dhc;
    const float uu = U[cell], c = C[cell];
    const float hp = h_prev_of(HS, td, dir, T, B, H, b, u);
    duv = dht * (hp - c);
    dzc = dht * (1.f - uu) * (1.f - c * c);
    dd = dht * uu;
  }
  dh[i] = dhc;                 // kept for inactive frames (carried through)
  du[i] = duv; dhd[i] = dd; dzc_s[i] = dzc; dzc_all[cell] = dzc;
}

// ---- BPTT, stage 2: through r*h and the gate sigmoids
__global__ void __launch_bounds__(256)
gru_bwd_gates_kernel(int s, int T, int B, int H, const float* __restrict__ R, const float* __restrict__ U,
                     const float* __restrict__ HS, const int* __restrict__ seq_len, const float* __restrict__ drh,
                     const float* __restrict__ du, const float* __restrict__ dhd, float* __restrict__ dh,
                     float* __restrict__ dzru_s, float* __restrict__ dzru_all, float* __restrict__ rh_all) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * H) return;
  const int u = i % H, b = (i / H) % B, dir = i / (H * B);
  const int td = frame_of(T - 1 - s, dir, T);
  const size_t row = (size_t)td * B + b;
  const size_t cell = (row * 2 + dir) * H + u;
  const bool active = td < seq_len[b];
  float dzr = 0.f, dzu = 0.f, rhv = 0.f;
  if (active) {
    const float r = R[cell], uu = U[cell];
    const float hp = h_prev_of(HS, td, dir, T, B, H, b, u);
    const float d = drh[i];
    dzr = d * hp * r * (1.f - r);
    dzu = du[i] * uu * (1.f - uu);
    dh[i] = dhd[i] + d * r;      // + dz_gates . U_gates^T, added by the next stage-1 launch
    rhv = r * hp;
  }
  float* zs = dzru_s + ((size_t)dir * B + b) * 2 * H;
  zs[u] = dzr; zs[H + u] = dzu;
  float* za = dzru_all + (row * 2 + dir) * 2 * H;
  za[u] = dzr; za[H + u] = dzu;
  rh_all[cell] = rhv;
}

__global__ void __launch_bounds__(256)
gru_final_state_kernel(int n, const float* __restrict__ hstate, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = hstate[i];
}

int gru_check(const b2_gru_desc* d) {
  B2_CHECK_ARG(d, "gru: null descriptor");
  B2_CHECK_ARG(d->T > 0 && d->B > 0 && d->D_in > 0 && d->H > 0, "gru: bad shape T=%d B=%d D=%d H=%d", d->T, d->B,
               d->D_in, d->H);
  B2_CHECK_ARG(d->B <= 64, "gru: batch %d > 64 (skinny recurrent products)", d->B);
  B2_CHECK_ARG(d->keep_prob > 0.f && d->keep_prob <= 1.f, "gru: keep_prob %f", d->keep_prob);
  return B2_OK;
}

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" size_t b2_bgru_reserve_bytes(const b2_gru_desc* d) { return d ? gru_reserve_layout(d, nullptr, nullptr) : 0; }
extern "C" size_t b2_bgru_workspace_bytes(const b2_gru_desc* d) { return d ? gru_work_layout(d, nullptr, nullptr) : 0; }

extern "C" int b2_bgru_layer_forward(const b2_gru_desc* d, const float* x, const int32_t* seq_len,
                                     const b2_gru_params* fw, const b2_gru_params* bw, float* y, float* final_state,
                                     void* reserve, void* workspace, size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = gru_check(d);
  if (rc) return rc;
  B2_CHECK_ARG(x && seq_len && fw && bw && y && reserve && workspace, "bgru_forward: null pointer");
  GruWork w;
  if (workspace_bytes < gru_work_layout(d, workspace, &w)) { set_error("bgru_forward: workspace too small"); return B2_ERR_WORKSPACE; }
  GruReserve r;
  gru_reserve_layout(d, reserve, &r);
  const int T = d->T, B = d->B, D = d->D_in, H = d->H, TB = T * B;
  const b2_gru_params* P[2] = {fw, bw};
  // time-batched input projections (rows [0, D) of the TF kernels), bias included
  for (int dir = 0; dir < 2; ++dir) {
    rc = gemm_simt(0, 0, TB, 2 * H, D, 1.f, x, D, P[dir]->gates_kernel, 2 * H, 0.f, w.gru + (size_t)dir * 2 * H, 4 * H,
                   P[dir]->gates_bias, stream);
    if (rc) return rc;
    rc = gemm_simt(0, 0, TB, H, D, 1.f, x, D, P[dir]->cand_kernel, H, 0.f, w.gc + (size_t)dir * H, 2 * H,
                   P[dir]->cand_bias, stream);
    if (rc) return rc;
  }
  B2_CUDA(cudaMemsetAsync(w.hstate, 0, (size_t)2 * B * H * sizeof(float), stream));
  const int n = 2 * B * H, blocks = cdiv(n, 256);
  const float* Ug[2] = {fw->gates_kernel + (size_t)D * 2 * H, bw->gates_kernel + (size_t)D * 2 * H};
  const float* Uc[2] = {fw->cand_kernel + (size_t)D * H, bw->cand_kernel + (size_t)D * H};
  for (int s = 0; s < T; ++s) {
    rc = gemm_skinny_pair(0, B, 2 * H, H, w.hstate, w.hstate + (size_t)B * H, H, Ug[0], Ug[1], 2 * H, w.zru,
                          w.zru + (size_t)B * 2 * H, 2 * H, stream);
    if (rc) return rc;
    gru_fwd_gates_kernel<<<blocks, 256, 0, stream>>>(s, T, B, H, w.gru, w.zru, w.hstate, r.r, r.u, w.rh);
    rc = gemm_skinny_pair(0, B, H, H, w.rh, w.rh + (size_t)B * H, H, Uc[0], Uc[1], H, w.zc, w.zc + (size_t)B * H, H, stream);
    if (rc) return rc;
    gru_fwd_out_kernel<<<blocks, 256, 0, stream>>>(s, T, B, H, w.gc, w.zc, r.u, seq_len, d->keep_prob, d->dropout_seed,
                                                   w.hstate, r.c, r.hs, y);
  }
  count_launches(2 * T);
  B2_CUDA(cudaGetLastError());
  if (final_state) {
    gru_final_state_kernel<<<blocks, 256, 0, stream>>>(n, w.hstate, final_state);
    B2_LAUNCH_CHECK();
  }
  return B2_OK;
}

extern "C" int b2_bgru_layer_backward(const b2_gru_desc* d, const float* x, const int32_t* seq_len,
                                      const b2_gru_params* fw, const b2_gru_params* bw, const float* dy,
                                      const void* reserve, float* dx, const b2_gru_grads* g_fw,
                                      const b2_gru_grads* g_bw, void* workspace, size_t workspace_bytes,
                                      b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = gru_check(d);
  if (rc) return rc;
  B2_CHECK_ARG(x && seq_len && fw && bw && dy && reserve && g_fw && g_bw && workspace, "bgru_backward: null pointer");
  GruWork w;
  if (workspace_bytes < gru_work_layout(d, workspace, &w)) { set_error("bgru_backward: workspace too small"); return B2_ERR_WORKSPACE; }
  GruReserve r;
  gru_reserve_layout(d, (void*)reserve, &r);
  const int T = d->T, B = d->B, D = d->D_in, H = d->H, TB = T * B;
  const b2_gru_params* P[2] = {fw, bw};
  const b2_gru_grads* G[2] = {g_fw, g_bw};
  const int n = 2 * B * H, blocks = cdiv(n, 256);
  const float* Ug[2] = {fw->gates_kernel + (size_t)D * 2 * H, bw->gates_kernel + (size_t)D * 2 * H};
  const float* Uc[2] = {fw->cand_kernel + (size_t)D * H, bw->cand_kernel + (size_t)D * H};
  B2_CUDA(cudaMemsetAsync(w.hstate, 0, (size_t)n * sizeof(float), stream));     // carried dh
  for (int s = 0; s < T; ++s) {
    gru_bwd_out_kernel<<<blocks, 256, 0, stream>>>(s, T, B, H, dy, r.u, r.c, r.hs, seq_len, d->keep_prob,
                                                   d->dropout_seed, s == 0, w.dhp, w.hstate, w.du, w.dhd, w.rh, w.gc);
    // d(r*h) = dz_c . U_cand^T
    rc = gemm_skinny_pair(1, B, H, H, w.rh, w.rh + (size_t)B * H, H, Uc[0], Uc[1], H, w.zc, w.zc + (size_t)B * H, H, stream);
    if (rc) return rc;
    gru_bwd_gates_kernel<<<blocks, 256, 0, stream>>>(s, T, B, H, r.r, r.u, r.hs, seq_len, w.zc, w.du, w.dhd, w.hstate,
                                                     w.zru, w.gru, w.rhs);
    // dh (into the previous state) += dz_gates . U_gates^T: added by the next stage-1 launch
    rc = gemm_skinny_pair(1, B, H, 2 * H, w.zru, w.zru + (size_t)B * 2 * H, 2 * H, Ug[0], Ug[1], 2 * H, w.dhp,
                          w.dhp + (size_t)B * H, H, stream);
    if (rc) return rc;
  }
  count_launches(2 * T);
  B2_CUDA(cudaGetLastError());
  // time-batched: weight, bias and input gradients.  h_prev of frame t = HS of the frame visited before it: a one-frame
  // shift of the [T*B] rows (fw: t-1, bw: t+1); the first frame's h_prev is 0 and contributes nothing.
  for (int dir = 0; dir < 2; ++dir) {
    const float* dzg = w.gru + (size_t)dir * 2 * H;       // [TB, 2H] pitch 4H
    const float* dzc = w.gc + (size_t)dir * H;            // [TB, H]  pitch 2H
    rc = b2_colsum(dzg, TB, 2 * H, 4 * H, G[dir]->gates_bias, 1, stream_);
    if (rc) return rc;
    rc = b2_colsum(dzc, TB, H, 2 * H, G[dir]->cand_bias, 1, stream_);
    if (rc) return rc;
    rc = gemm_simt(1, 0, D, 2 * H, TB, 1.f, x, D, dzg, 4 * H, 1.f, G[dir]->gates_kernel, 2 * H, nullptr, stream);
    if (rc) return rc;
    rc = gemm_simt(1, 0, D, H, TB, 1.f, x, D, dzc, 2 * H, 1.f, G[dir]->cand_kernel, H, nullptr, stream);
    if (rc) return rc;
    if (T > 1) {
      const float* hs_a = r.hs + (size_t)dir * H + (dir == 0 ? 0 : (size_t)B * 2 * H);
      const float* dzg_h = dzg + (dir == 0 ? (size_t)B * 4 * H : 0);
      rc = gemm_simt(1, 0, H, 2 * H, (T - 1) * B, 1.f, hs_a, 2 * H, dzg_h, 4 * H, 1.f,
                     G[dir]->gates_kernel + (size_t)D * 2 * H, 2 * H, nullptr, stream);
      if (rc) return rc;
    }
    // candidate recurrent weights: (r * h_prev)^T . dz_c, r*h_prev stored per frame by the BPTT kernels
    rc = gemm_simt(1, 0, H, H, TB, 1.f, w.rhs + (size_t)dir * H, 2 * H, dzc, 2 * H, 1.f,
                   G[dir]->cand_kernel + (size_t)D * H, H, nullptr, stream);
    if (rc) return rc;
    if (dx) {
      rc = gemm_simt(0, 1, TB, D, 2 * H, 1.f, dzg, 4 * H, P[dir]->gates_kernel, 2 * H, dir == 0 ? 0.f : 1.f, dx, D,
                     nullptr, stream);
      if (rc) return rc;
      rc = gemm_simt(0, 1, TB, D, H, 1.f, dzc, 2 * H, P[dir]->cand_kernel, H, 1.f, dx, D, nullptr, stream);
      if (rc) return rc;
    }
  }
  return B2_OK;
}
