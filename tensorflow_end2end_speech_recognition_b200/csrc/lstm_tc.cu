// bf16 / tcgen05 BLSTM layer: orchestration of the time-batched GEMMs (gemm_tcgen05.cu)
// around the persistent recurrence kernels (lstm_rec_tc.cu), plus the weight packing that
// puts every operand in the order those kernels want.
//
// Packed gate order: column n' = dir*4H + u*4 + gate (unit-major, gate-minor) instead of
// TensorFlow's dir / gate*H + u (models/encoders/core/blstm.py:287-305 via LSTMBlockCell):
// a CTA that owns 32 units then reads/writes one contiguous 128-float segment per frame.
// The permutation is applied once per step to the (small) weights; activations, gate
// pre-activations G and gate gradients dG live only in packed order.
#include "lstm_rec_tc.cuh"

namespace b2 {

enum { EPI_STORE_F32 = 0, EPI_ATOMIC_F32 = 1, EPI_STORE_BF16 = 2 };

static size_t pad8(size_t x) { return (x + 7) / 8 * 8; }

struct TcWork {
  float* G;                  // [TB, 8H] fp32 gate pre-activations (forward)
  __nv_bfloat16* dG;         // [TB, 8H] bf16 gate gradients (backward), set 0 / set 1 below
  __nv_bfloat16* dG2[2]; float* dwx2[2]; float* dwh2[2]; float* dbias2[2];
  __nv_bfloat16* xb;         // [TB, pad8(D)] bf16 copy of x when the caller has none
  __nv_bfloat16* wx;         // [D, 8H] packed input weights
  float* bias;               // [8H] packed bias
  uint16_t* wh;              // forward recurrent pack  [2][CS][128][H]
  uint16_t* whT;             // backward recurrent pack [2][CS][4][128][128]
  float* dwx;                // [D, 8H] fp32 packed weight gradient (scratch)
  float* dwh;                // [2][H, 4H] fp32 packed recurrent weight gradient (scratch)
  float* dbias;              // [8H]
  float* dym;                // [TB, 2H] dropout-masked dy (backward, keep_prob < 1 only)
};

static size_t tc_work_layout(const b2_lstm_desc* d, void* base, TcWork* w) {
  const size_t TB = (size_t)d->T * d->B, H = d->H, D = d->D_in;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  // G (forward) and dG (backward) are never live together
  const size_t oG = take(TB * 8 * H * 4);
  const size_t oxb = take(TB * pad8(D) * 2);
  const size_t owx = take(D * 8 * H * 2);
  const size_t ob = take(8 * H * 4);
  const size_t owh = take(2 * 4 * H * H * 2);
  const size_t owt = take((size_t)2 * (H / 32) * 4 * 128 * 128 * 2);   // whT: 4 zero-padded M tiles
  const size_t odwx = take(D * 8 * H * 4);
  const size_t odwh = take(2 * H * 4 * H * 4);
  const size_t odb = take(8 * H * 4);
  // second set of backward buffers: the weight-gradient GEMMs of layer l run on a side stream
  // while layer l-1's recurrence already fills the other set
  const size_t oG1 = take(TB * 8 * H * 2);
  const size_t odwx1 = take(D * 8 * H * 4);
  const size_t odwh1 = take(2 * H * 4 * H * 4);
  const size_t odb1 = take(8 * H * 4);
  const size_t odym = d->keep_prob < 1.f ? take(TB * 2 * H * 4) : 0;
  if (w) {
    char* p = (char*)base;
    w->G = (float*)(p + oG); w->dG = (__nv_bfloat16*)(p + oG); w->xb = (__nv_bfloat16*)(p + oxb);
    w->wx = (__nv_bfloat16*)(p + owx); w->bias = (float*)(p + ob); w->wh = (uint16_t*)(p + owh);
    w->whT = (uint16_t*)(p + owt); w->dwx = (float*)(p + odwx); w->dwh = (float*)(p + odwh);
    w->dbias = (float*)(p + odb);
    w->dG2[0] = w->dG; w->dwx2[0] = w->dwx; w->dwh2[0] = w->dwh; w->dbias2[0] = w->dbias;
    w->dG2[1] = (__nv_bfloat16*)(p + oG1); w->dwx2[1] = (float*)(p + odwx1);
    w->dwh2[1] = (float*)(p + odwh1); w->dbias2[1] = (float*)(p + odb1);
    w->dym = d->keep_prob < 1.f ? (float*)(p + odym) : nullptr;
  }
  return off;
}

// DropoutWrapper backward, hoisted out of the BPTT kernel: dy_masked[i] = keep(seed, i) ? dy[i]/keep : 0 over the
// [T,B,2H] layout (the element index IS the mask counter).  Inside the kernel the hash sat on the dependent chain of the
// gate-math warps (+0.29 ms per layer at config 2); as a streaming pass it costs the HBM time of 2 x 262 MB.
__global__ void __launch_bounds__(256)
dropout_mask_dy_kernel(const float* __restrict__ dy, float* __restrict__ out, int64_t n4, float keep,
                       unsigned long long seed) {
  const float sc = 1.f / keep;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = ((const float4*)dy)[i];
    const uint64_t e = (uint64_t)i * 4;
    v.x = dropout_keep(seed, e, keep) ? v.x * sc : 0.f;
    v.y = dropout_keep(seed, e + 1, keep) ? v.y * sc : 0.f;
    v.z = dropout_keep(seed, e + 2, keep) ? v.z * sc : 0.f;
    v.w = dropout_keep(seed, e + 3, keep) ? v.w * sc : 0.f;
    ((float4*)out)[i] = v;
  }
}

// ---------------------------------------------------------------- packing kernels
// kernel_dir [(D+H), 4H] fp32 (TF layout) ->
//   wx  [D, 8H] bf16   wx[k][dir*4H + u*4 + g]          = kernel_dir[k][g*H + u]
//   bias[8H]    fp32   bias[dir*4H + u*4 + g]           = bias_dir[g*H + u]
//   wh  [2][CS][128][H]      wh[dir][cta][ul*4+g][k]    = kernel_dir[D + k][g*H + cta*32 + ul]
//   whT [2][CS][4][128][128] whT[dir][cta][m][i][ul*4+g] = kernel_dir[D + 128m + i][g*H + cta*32 + ul]
__global__ void pack_lstm_weights_kernel(const float* __restrict__ k0, const float* __restrict__ k1,
                                         const float* __restrict__ b0, const float* __restrict__ b1,
                                         int D, int H, __nv_bfloat16* __restrict__ wx,
                                         float* __restrict__ bias, uint16_t* __restrict__ wh,
                                         uint16_t* __restrict__ whT) {
  const int64_t n_wx = (int64_t)D * 8 * H, n_wh = (int64_t)2 * 4 * H * H;
  const int64_t n_wt = (int64_t)2 * (H / 32) * 4 * 128 * 128;
  const int64_t total = n_wx + 8 * H + n_wh + (whT ? n_wt : 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n_wx) {
      const int col = (int)(i % (8 * H)); const int k = (int)(i / (8 * H));
      const int dir = col / (4 * H), r = col % (4 * H), u = r >> 2, g = r & 3;
      wx[i] = __float2bfloat16((dir ? k1 : k0)[(size_t)k * 4 * H + g * H + u]);
    } else if (i < n_wx + 8 * H) {
      const int col = (int)(i - n_wx);
      const int dir = col / (4 * H), r = col % (4 * H), u = r >> 2, g = r & 3;
      bias[col] = (dir ? b1 : b0)[g * H + u];
    } else if (i < n_wx + 8 * H + n_wh) {
      const int64_t j = i - n_wx - 8 * H;
      const int k = (int)(j % H); const int64_t rr = j / H;          // rr = dir*4H + cta*128 + r
      const int dir = (int)(rr / (4 * H)); const int q = (int)(rr % (4 * H));
      const int cta = q / 128, r = q % 128, ul = r >> 2, g = r & 3;
      const float v = (dir ? k1 : k0)[(size_t)(D + k) * 4 * H + g * H + cta * 32 + ul];
      wh[j] = __bfloat16_as_ushort(__float2bfloat16(v));
    } else {
      const int64_t j = i - n_wx - 8 * H - n_wh;
      const int r = (int)(j % 128); const int64_t t1 = j / 128;
      const int row = (int)(t1 % 128); const int64_t t2 = t1 / 128;
      const int m = (int)(t2 % 4); const int64_t t3 = t2 / 4;
      const int CS = H / 32;
      const int cta = (int)(t3 % CS), dir = (int)(t3 / CS);
      const int ul = r >> 2, g = r & 3;
      float v = 0.f;
      if (128 * m + row < H)
        v = (dir ? k1 : k0)[(size_t)(D + 128 * m + row) * 4 * H + g * H + cta * 32 + ul];
      whT[j] = __bfloat16_as_ushort(__float2bfloat16(v));
    }
  }
}

// scatter-add packed gradients back into the TF layout:
//   gk_dir[k][g*H+u]      += dwx[k][dir*4H + u*4 + g]              k < D
//   gk_dir[D+k][g*H+u]    += dwh[dir][k][u*4 + g]
//   gb_dir[g*H+u]         += dbias[dir*4H + u*4 + g]
__global__ void unpack_lstm_grads_kernel(const float* __restrict__ dwx, const float* __restrict__ dwh,
                                         const float* __restrict__ dbias, int D, int H,
                                         float* __restrict__ gk0, float* __restrict__ gk1,
                                         float* __restrict__ gb0, float* __restrict__ gb1) {
  const int64_t n_wx = (int64_t)D * 8 * H, n_wh = (int64_t)2 * H * 4 * H;
  const int64_t total = n_wx + n_wh + 8 * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n_wx) {
      const int col = (int)(i % (8 * H)); const int k = (int)(i / (8 * H));
      const int dir = col / (4 * H), r = col % (4 * H), u = r >> 2, g = r & 3;
      (dir ? gk1 : gk0)[(size_t)k * 4 * H + g * H + u] += dwx[i];
    } else if (i < n_wx + n_wh) {
      const int64_t j = i - n_wx;
      const int r = (int)(j % (4 * H)); const int64_t t1 = j / (4 * H);
      const int k = (int)(t1 % H), dir = (int)(t1 / H);
      const int u = r >> 2, g = r & 3;
      (dir ? gk1 : gk0)[(size_t)(D + k) * 4 * H + g * H + u] += dwh[j];
    } else {
      const int col = (int)(i - n_wx - n_wh);
      const int dir = col / (4 * H), r = col % (4 * H), u = r >> 2, g = r & 3;
      (dir ? gb1 : gb0)[g * H + u] += dbias[col];
    }
  }
}

// ------------------------------------------------------------------ kernel timers
// Optional CUDA-event timers around the recurrence launches (bench.py's live roofline).
static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // fwd0 fwd1 bwd0 bwd1
static void prof_record(int idx, cudaStream_t stream) {
  if (!g_prof_on) return;
  if (!g_prof_ev[0]) for (int i = 0; i < 4; ++i) cudaEventCreate(&g_prof_ev[i]);
  cudaEventRecord(g_prof_ev[idx], stream);
}
void tc_profile_enable(int on) { g_prof_on = on != 0; }
int tc_profile_last_ms(float* fwd_ms, float* bwd_ms) {
  if (!g_prof_ev[0]) return B2_ERR_INVALID;
  if (fwd_ms && cudaEventElapsedTime(fwd_ms, g_prof_ev[0], g_prof_ev[1]) != cudaSuccess) *fwd_ms = -1.f;
  if (bwd_ms && cudaEventElapsedTime(bwd_ms, g_prof_ev[2], g_prof_ev[3]) != cudaSuccess) *bwd_ms = -1.f;
  return B2_OK;
}

// ------------------------------------------------------------------ side stream
// Weight-gradient GEMMs are off the BPTT critical path: they run on a low-priority side
// stream, capped to the SMs the recurrence clusters leave free, while the next layer's
// recurrence proceeds on the caller's stream.
struct SideCtx {
  cudaStream_t s = nullptr;
  cudaEvent_t ev_main = nullptr, ev_done[2] = {nullptr, nullptr};
  bool pending[2] = {false, false};
  int toggle = 0;
};
static SideCtx g_side[16];
static SideCtx* side_ctx() {
  int dev = 0;
  cudaGetDevice(&dev);
  SideCtx* c = &g_side[dev & 15];
  if (!c->s) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&c->s, cudaStreamNonBlocking, lo);
    cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_done[1], cudaEventDisableTiming);
  }
  return c;
}
// make `stream` wait for every outstanding side-stream gradient GEMM
int tc_backward_join(cudaStream_t stream) {
  SideCtx* c = side_ctx();
  for (int k = 0; k < 2; ++k)
    if (c->pending[k]) {
      B2_CUDA(cudaStreamWaitEvent(stream, c->ev_done[k], 0));
      c->pending[k] = false;
    }
  return B2_OK;
}

// ------------------------------------------------------------------ layer entry points
bool tc_layer_supported(const b2_lstm_desc* d) {
  static int sm100 = -1;
  if (d->precision != B2_PREC_BF16 || !rec_tc_supported(d->H) || d->num_proj > 0) return false;
  if (!env_int("B2_REC_TC", 1)) return false;
  if (sm100 < 0) sm100 = b2_device_is_sm100();
  return sm100 == 1;
}

size_t tc_layer_workspace_bytes(const b2_lstm_desc* d) { return tc_work_layout(d, nullptr, nullptr); }

static int pack_weights(const b2_lstm_desc* d, const b2_lstm_params* fw, const b2_lstm_params* bw,
                        const TcWork& w, bool need_T, cudaStream_t stream) {
  pack_lstm_weights_kernel<<<num_sms() * 8, 256, 0, stream>>>(
      fw->kernel, bw->kernel, fw->bias, bw->bias, d->D_in, d->H, w.wx, w.bias, w.wh,
      need_T ? w.whT : nullptr);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

int tc_layer_forward(const b2_lstm_desc* d, const float* x, const __nv_bfloat16* x_lp,
                     const int32_t* seq_len, const b2_lstm_params* fw, const b2_lstm_params* bw,
                     float* y, float* final_state, void* reserve, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream) {
  TcWork w;
  const size_t need = tc_work_layout(d, workspace, &w);
  if (workspace_bytes < need) { set_error("tc_layer_forward: workspace %zu < %zu", workspace_bytes, need); return B2_ERR_WORKSPACE; }
  Reserve r;
  reserve_layout(d, reserve, &r);
  const int T = d->T, B = d->B, D = d->D_in, H = d->H, TB = T * B;
  int rc = tc_backward_join(stream);
  if (rc) return rc;
  rc = pack_weights(d, fw, bw, w, false, stream);
  if (rc) return rc;
  const __nv_bfloat16* xa = x_lp;
  int ldx = D;
  if (!xa || (D % 8)) {
    ldx = (int)pad8(D);
    rc = cast_f32_bf16(x, TB, D, D, w.xb, ldx, stream);
    if (rc) return rc;
    xa = w.xb;
  }
  // G[TB, 8H] = X . Wx_packed + bias_packed   (both directions in one GEMM)
  rc = gemm_bf16_tc(0, 1, TB, 8 * H, D, 1.f, xa, ldx, w.wx, 8 * H, w.G, 8 * H, w.bias,
                    EPI_STORE_F32, 0, stream);
  if (rc) return rc;
  RecFwdArgs ra;
  ra.T = T; ra.B = B; ra.H = H; ra.NG = 0; ra.seq_len = seq_len; ra.wpack = w.wh;
  const b2_lstm_params* P[2] = {fw, bw};
  for (int dir = 0; dir < 2; ++dir) {
    ra.wi[dir] = P[dir]->w_i_diag; ra.wf[dir] = P[dir]->w_f_diag; ra.wo[dir] = P[dir]->w_o_diag;
  }
  ra.use_peephole = d->use_peephole; ra.forget_bias = d->forget_bias; ra.cell_clip = d->cell_clip;
  ra.keep_prob = d->keep_prob; ra.seed = d->dropout_seed;
  ra.y = y; ra.hs_lp = r.hs_lp; ra.y_lp = r.y_lp;
  ra.gates = d->need_backward ? r.gates : nullptr; ra.cs = d->need_backward ? r.cs : nullptr;
  ra.final_state = final_state; ra.dbg = nullptr;
  if (env_int("B2_REC_DBG", 0)) {
    static long long* dbg_buf = nullptr;
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 64 * sizeof(long long), stream);
    ra.dbg = dbg_buf;
    rc = rec_tc_forward(ra, w.G, env_int("B2_REC_NCHAIN", 0), stream);
    long long hb[32];
    cudaMemcpyAsync(hb, dbg_buf, sizeof(hb), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    fprintf(stderr, "[rec fwd dbg3] loop cycles per (cluster,chain): %lld %lld | %lld %lld | %lld %lld | %lld %lld\n",
            hb[16], hb[17], hb[18], hb[19], hb[20], hb[21], hb[22], hb[23]);
    fprintf(stderr, "[rec fwd dbg] cycles/step: mma_wait_h=%lld mma_issue=%lld | epi wait_acc=%lld "
            "ld+transpose=%lld wait_G=%lld math+saves=%lld fence+bar=%lld send+store=%lld\n",
            hb[0] / T, hb[1] / T, hb[2] / T, hb[3] / T, hb[4] / T, hb[5] / T, hb[6] / T, hb[7] / T);
    return rc;
  }
  prof_record(0, stream);
  rc = rec_tc_forward(ra, w.G, env_int("B2_REC_NCHAIN", 0), stream);
  prof_record(1, stream);
  return rc;
}

int tc_layer_backward(const b2_lstm_desc* d, const float* x, const __nv_bfloat16* x_lp,
                      const int32_t* seq_len, const b2_lstm_params* fw, const b2_lstm_params* bw,
                      const float* dy, const float* d_final_state, const void* reserve, float* dx,
                      const b2_lstm_grads* g_fw,
                      const b2_lstm_grads* g_bw, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
  TcWork w;
  const size_t need = tc_work_layout(d, workspace, &w);
  if (workspace_bytes < need) { set_error("tc_layer_backward: workspace %zu < %zu", workspace_bytes, need); return B2_ERR_WORKSPACE; }
  Reserve r;
  reserve_layout(d, (void*)reserve, &r);
  const int T = d->T, B = d->B, D = d->D_in, H = d->H, TB = T * B;
  // measured on B200: co-scheduling the GEMM CTAs delays the 16-CTA clusters of the next recurrence
  // (41.7 vs 35.6 ms/step), so the overlap is opt-in
  const bool use_side = env_int("B2_SIDE_STREAM", 0) != 0;
  SideCtx* sc = side_ctx();
  const int k = use_side ? sc->toggle : 0;
  if (sc->pending[k]) {            // the side work that last used buffer set k must be done
    B2_CUDA(cudaStreamWaitEvent(stream, sc->ev_done[k], 0));
    sc->pending[k] = false;
  }
  __nv_bfloat16* dG = w.dG2[k];
  float* dwx = w.dwx2[k]; float* dwh = w.dwh2[k]; float* dbias = w.dbias2[k];
  int rc = pack_weights(d, fw, bw, w, true, stream);
  if (rc) return rc;
  // 1. BPTT recurrence -> dG (bf16, packed order); bias / peephole gradients accumulate in
  //    registers inside the kernel and are flushed with atomics
  RecBwdArgs ba;
  ba.T = T; ba.B = B; ba.H = H; ba.NG = 0; ba.seq_len = seq_len; ba.wpackT = w.whT;
  const b2_lstm_params* P[2] = {fw, bw};
  const b2_lstm_grads* GR[2] = {g_fw, g_bw};
  for (int dir = 0; dir < 2; ++dir) {
    ba.wi[dir] = P[dir]->w_i_diag; ba.wf[dir] = P[dir]->w_f_diag; ba.wo[dir] = P[dir]->w_o_diag;
    ba.dwi[dir] = GR[dir]->w_i_diag; ba.dwf[dir] = GR[dir]->w_f_diag; ba.dwo[dir] = GR[dir]->w_o_diag;
  }
  ba.use_peephole = d->use_peephole; ba.cell_clip = d->cell_clip; ba.keep_prob = d->keep_prob;
  ba.seed = d->dropout_seed; ba.gates = r.gates; ba.cs = r.cs; ba.dG = dG; ba.dfinal = d_final_state;
  if (d->keep_prob < 1.f && (((size_t)TB * 2 * H) % 4 == 0)) {     // mask dy once, outside the recurrence
    const int64_t n4 = (int64_t)TB * 2 * H / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    dropout_mask_dy_kernel<<<(int)blocks, 256, 0, stream>>>(dy, w.dym, n4, d->keep_prob, d->dropout_seed);
    B2_LAUNCH_CHECK();
    dy = w.dym;
    ba.keep_prob = 1.f;
  }
  B2_CUDA(cudaMemsetAsync(dbias, 0, (size_t)8 * H * 4, stream));
  ba.dbias = dbias;
  ba.dbg = nullptr;
  const int nchain = env_int("B2_REC_NCHAIN", 0);
  if (env_int("B2_REC_DBG", 0)) {
    static long long* dbg_buf = nullptr;
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 64 * sizeof(long long), stream);
    ba.dbg = dbg_buf;
    rc = rec_tc_backward(ba, dy, nchain, stream);
    long long hb[12];
    cudaMemcpyAsync(hb, dbg_buf, sizeof(hb), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    fprintf(stderr, "[rec bwd dbg2] loads=%lld math=%lld stores=%lld\n", hb[8] / T, hb[9] / T, hb[10] / T);
    fprintf(stderr, "[rec bwd dbg] cycles/step: wait_partials+sum=%lld wait_ring=%lld math+stores=%lld "
            "bar1=%lld wait_mma=%lld ld+convert=%lld bar2+send=%lld\n",
            hb[0] / T, hb[1] / T, hb[2] / T, hb[3] / T, hb[4] / T, hb[5] / T, hb[6] / T);
  } else {
    prof_record(2, stream);
    rc = rec_tc_backward(ba, dy, nchain, stream);
    prof_record(3, stream);
  }
  if (rc) return rc;
  // 2. critical path: dX[TB, D] = dG[TB, 8H] . Wx_packed^T  (sums both directions)
  const __nv_bfloat16* xa = x_lp;
  int ldx = D;
  if (!xa || (D % 8)) {
    ldx = (int)pad8(D);
    rc = cast_f32_bf16(x, TB, D, D, w.xb, ldx, stream);
    if (rc) return rc;
    xa = w.xb;
  }
  if (dx) {
    rc = gemm_bf16_tc(0, 0, TB, D, 8 * H, 1.f, dG, 8 * H, w.wx, 8 * H, dx, D, nullptr,
                      EPI_STORE_F32, 0, stream);
    if (rc) return rc;
  }
  // 3. off the critical path: weight gradients on packed operands
  cudaStream_t gs = stream;
  if (use_side) {
    B2_CUDA(cudaEventRecord(sc->ev_main, stream));
    B2_CUDA(cudaStreamWaitEvent(sc->s, sc->ev_main, 0));
    gs = sc->s;
    // leave the SMs of the next layer's recurrence clusters alone
    const int ng = cdiv(B, RN);
    const int nch = nchain > 0 ? (nchain > 2 ? 2 : nchain) : (ng >= 2 ? 2 : 1);
    const int rec_ctas = 2 * cdiv(ng, nch) * (H / RU);
    int free_sms = num_sms() - rec_ctas;
    if (free_sms < 16) free_sms = 16;
    gemm_set_cta_limit(free_sms);
  }
  // dWx_packed[D, 8H] = X^T . dG
  B2_CUDA(cudaMemsetAsync(dwx, 0, (size_t)D * 8 * H * 4, gs));
  rc = gemm_bf16_tc(1, 1, D, 8 * H, TB, 1.f, xa, ldx, dG, 8 * H, dwx, 8 * H, nullptr,
                    EPI_ATOMIC_F32, 0, gs);
  // dWh_packed[dir][H, 4H] = Hprev_dir^T . dG_dir   (hs shifted by one step)
  if (!rc) rc = (cudaMemsetAsync(dwh, 0, (size_t)2 * H * 4 * H * 4, gs) == cudaSuccess) ? B2_OK : B2_ERR_CUDA;
  if (!rc && T > 1) {
    for (int dir = 0; dir < 2 && !rc; ++dir) {
      const __nv_bfloat16* ha = r.hs_lp + (size_t)dir * H + (dir == 0 ? 0 : (size_t)B * 2 * H);
      const __nv_bfloat16* gb = dG + (size_t)dir * 4 * H + (dir == 0 ? (size_t)B * 8 * H : 0);
      rc = gemm_bf16_tc(1, 1, H, 4 * H, (T - 1) * B, 1.f, ha, 2 * H, gb, 8 * H,
                        dwh + (size_t)dir * H * 4 * H, 4 * H, nullptr, EPI_ATOMIC_F32, 0, gs);
    }
  }
  gemm_set_cta_limit(0);
  if (rc) return rc;
  unpack_lstm_grads_kernel<<<num_sms() * 4, 256, 0, gs>>>(dwx, dwh, dbias, D, H, g_fw->kernel,
                                                        g_bw->kernel, g_fw->bias, g_bw->bias);
  B2_LAUNCH_CHECK();
  if (use_side) {
    B2_CUDA(cudaEventRecord(sc->ev_done[k], sc->s));
    sc->pending[k] = true;
    sc->toggle ^= 1;
    // the first layer is the last one of a backward pass: hand the gradients back
    if (!dx) return tc_backward_join(stream);
  }
  return B2_OK;
}

}  // namespace b2
