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
  float* G2;                 // second buffer: consecutive layers alternate, so that layer l+1's chunked gate GEMM can
                             // fill its G while layer l's recurrence still reads its own
  __nv_bfloat16* dG;         // [TB, 8H] bf16 gate gradients (backward; only when the reserve holds none)
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
  const size_t oG2 = take(TB * 8 * H * 4);
  const size_t oxb = take(TB * pad8(D) * 2);
  const size_t owx = take(D * 8 * H * 2);
  const size_t ob = take(8 * H * 4);
  const size_t owh = take(2 * 4 * H * H * 2);
  const size_t owt = take((size_t)2 * (H / 32) * 4 * 128 * 128 * 2);   // whT: 4 zero-padded M tiles
  const size_t odwx = take(D * 8 * H * 4);
  const size_t odwh = take(2 * H * 4 * H * 4);
  const size_t odb = take(8 * H * 4);
  const size_t odym = d->keep_prob < 1.f ? take(TB * 2 * H * 4) : 0;
  if (w) {
    char* p = (char*)base;
    w->G = (float*)(p + oG); w->G2 = (float*)(p + oG2); w->dG = (__nv_bfloat16*)(p + oG); w->xb = (__nv_bfloat16*)(p + oxb);
    w->wx = (__nv_bfloat16*)(p + owx); w->bias = (float*)(p + ob); w->wh = (uint16_t*)(p + owh);
    w->whT = (uint16_t*)(p + owt); w->dwx = (float*)(p + odwx); w->dwh = (float*)(p + odwh);
    w->dbias = (float*)(p + odb);
    w->dym = d->keep_prob < 1.f ? (float*)(p + odym) : nullptr;
  }
  return off;
}

// packed weights live in the layer's reserve when a backward pass will follow (one pack per layer and step)
static void use_reserve_pack(const b2_lstm_desc* d, void* pack, TcWork* w) {
  char* p = (char*)pack;
  const size_t H = d->H, D = d->D_in;
  w->wx = (__nv_bfloat16*)p;  p += align_up(D * 8 * H * 2, 1024);
  w->bias = (float*)p;        p += align_up(8 * H * 4, 1024);
  w->wh = (uint16_t*)p;       p += align_up(2 * 4 * H * H * 2, 1024);
  w->whT = (uint16_t*)p;
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
// Weight-gradient GEMMs are off the BPTT critical path.  With B2_SIDE_STREAM=1 (default) the GEMMs of layer l
// run on a low-priority side stream NEXT TO layer l-1's BPTT recurrence, on the SMs its clusters leave free:
//   * they are enqueued only after layer l-1's recurrence kernel has been launched, and the side stream
//     first waits (cuStreamWaitValue32) until every cluster of that kernel has reported itself resident --
//     otherwise the persistent GEMM CTAs spread over all GPCs and the 16-CTA clusters cannot be placed
//     until the GEMM drains (measured in round 1: 41.7 vs 35.6 ms/step);
//   * the GEMM grid is capped to the SMs the clusters do not use.
// The last layer of a backward pass has no recurrence to hide behind: its GEMMs are flushed by
// tc_backward_join.
struct DeferredWgrad {
  bool valid = false;
  int T = 0, B = 0, D = 0, H = 0, ldx = 0;
  const __nv_bfloat16* xa = nullptr; const __nv_bfloat16* dG = nullptr; const __nv_bfloat16* hs_lp = nullptr;
  float* dwx = nullptr; float* dwh = nullptr; const float* dbias = nullptr;
  float* gk0 = nullptr; float* gk1 = nullptr; float* gb0 = nullptr; float* gb1 = nullptr;
};
struct SideCtx {
  cudaStream_t s = nullptr;
  cudaEvent_t ev_rec = nullptr, ev_done = nullptr;
  bool pending = false;             // side-stream work whose completion `ev_done` marks
  DeferredWgrad def;
  unsigned* resident = nullptr;     // device counter: clusters of the recurrence kernels that are running
  unsigned expected = 0;            // its value once every cluster launched so far is resident
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
    cudaEventCreateWithFlags(&c->ev_rec, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming);
    cudaMalloc(&c->resident, sizeof(unsigned));
    cudaMemset(c->resident, 0, sizeof(unsigned));
  }
  return c;
}

typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, unsigned, unsigned);
static StreamWaitValue32Fn stream_wait_value_fn() {
  static StreamWaitValue32Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (StreamWaitValue32Fn)p;
  }
  return fn;
}

// dWx, dWh (packed order) and the scatter back into the TF layout, on stream gs
static int run_wgrad(const DeferredWgrad& w, cudaStream_t gs, int cta_limit) {
  const int TB = w.T * w.B, D = w.D, H = w.H;
  gemm_set_cta_limit(cta_limit);
  // dWx_packed[D, 8H] = X^T . dG
  int rc = (cudaMemsetAsync(w.dwx, 0, (size_t)D * 8 * H * 4, gs) == cudaSuccess) ? B2_OK : B2_ERR_CUDA;
  if (!rc) rc = gemm_bf16_tc(1, 1, D, 8 * H, TB, 1.f, w.xa, w.ldx, w.dG, 8 * H, w.dwx, 8 * H, nullptr,
                             EPI_ATOMIC_F32, 0, gs);
  // dWh_packed[dir][H, 4H] = Hprev_dir^T . dG_dir   (hs shifted by one step)
  if (!rc) rc = (cudaMemsetAsync(w.dwh, 0, (size_t)2 * H * 4 * H * 4, gs) == cudaSuccess) ? B2_OK : B2_ERR_CUDA;
  if (!rc && w.T > 1) {
    for (int dir = 0; dir < 2 && !rc; ++dir) {
      const __nv_bfloat16* ha = w.hs_lp + (size_t)dir * H + (dir == 0 ? 0 : (size_t)w.B * 2 * H);
      const __nv_bfloat16* gb = w.dG + (size_t)dir * 4 * H + (dir == 0 ? (size_t)w.B * 8 * H : 0);
      rc = gemm_bf16_tc(1, 1, H, 4 * H, (w.T - 1) * w.B, 1.f, ha, 2 * H, gb, 8 * H,
                        w.dwh + (size_t)dir * H * 4 * H, 4 * H, nullptr, EPI_ATOMIC_F32, 0, gs);
    }
  }
  gemm_set_cta_limit(0);
  if (rc) return rc;
  unpack_lstm_grads_kernel<<<num_sms() * 4, 256, 0, gs>>>(w.dwx, w.dwh, w.dbias, D, H, w.gk0, w.gk1, w.gb0, w.gb1);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// make `stream` wait for the side-stream work enqueued so far (NOT for the still-deferred weight gradients of the
// layer processed last): the per-layer gradient buckets of the data-parallel step are reduced behind this
int tc_backward_side_wait(cudaStream_t stream) {
  SideCtx* c = side_ctx();
  if (c->pending) B2_CUDA(cudaStreamWaitEvent(stream, c->ev_done, 0));
  return B2_OK;
}

// make `stream` wait for every outstanding side-stream gradient GEMM (flushing the deferred ones first)
int tc_backward_join(cudaStream_t stream) {
  SideCtx* c = side_ctx();
  if (c->def.valid) {
    B2_CUDA(cudaStreamWaitEvent(c->s, c->ev_rec, 0));        // dG of that layer is complete
    int rc = run_wgrad(c->def, c->s, 0);
    if (rc) return rc;
    B2_CUDA(cudaEventRecord(c->ev_done, c->s));
    c->pending = true;
    c->def.valid = false;
  }
  if (c->pending) {
    B2_CUDA(cudaStreamWaitEvent(stream, c->ev_done, 0));     // the side stream is in-order: covers all earlier work
    c->pending = false;
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

// ------------------------------------------------------------------ forward chunk chain
// Layer l+1's time-batched gate GEMM needs layer l's outputs of BOTH directions; frame chunk k is complete once the
// forward direction has passed it from below and the backward direction from above, i.e. the middle chunks first and
// all but the two outermost ones well before the recurrence kernel ends.  The recurrence kernel counts finished
// (CTA, chain) pairs per chunk (RecFwdArgs::progress); when the next layer's input IS that kernel's bf16 output, its
// gate GEMM is issued chunk by chunk on a second stream behind cuStreamWaitValue32 on those counters and runs on the
// SMs the clusters leave free, so that only the two outermost chunks remain on the critical path.
struct FwdChain {
  cudaStream_t s = nullptr;
  cudaEvent_t ev_launch = nullptr, ev_done = nullptr;
  unsigned* progress[2] = {nullptr, nullptr};   // [kMaxChunks] each, alternating between consecutive kernels
  int pp = 0;
  int gpar = 0;                                 // which G buffer the most recent layer used
  bool valid = false;
  const __nv_bfloat16* y_lp = nullptr;          // output of the most recent recurrence launch
  int T = 0, B = 0, cols = 0, chunk_T = 0, nchunks = 0, free_sms = 0;
  unsigned expected = 0;
};
constexpr int kMaxChunks = 64;
static FwdChain g_fwd_chain[16];
static FwdChain* fwd_chain() {
  int dev = 0;
  cudaGetDevice(&dev);
  FwdChain* c = &g_fwd_chain[dev & 15];
  if (!c->s) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&c->s, cudaStreamNonBlocking, hi);
    cudaEventCreateWithFlags(&c->ev_launch, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) {
      cudaMalloc(&c->progress[i], kMaxChunks * sizeof(unsigned));
      cudaMemset(c->progress[i], 0, kMaxChunks * sizeof(unsigned));
    }
  }
  return c;
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
  if (r.wpack) use_reserve_pack(d, r.wpack, &w);      // backward will reuse this pack (incl. the transposed slices)
  FwdChain* fc = fwd_chain();
  fc->gpar ^= 1;
  float* Gbuf = fc->gpar ? w.G2 : w.G;
  const int want_chunks = env_int("B2_FWD_CHUNKS", 16);
  const bool dbg_run = env_int("B2_REC_DBG", 0) != 0;
  // the chunk chain: this layer's input is the bf16 output of the recurrence launched last (still running, typically)
  const bool chained = want_chunks > 1 && !dbg_run && stream_wait_value_fn() != nullptr && r.wpack != nullptr &&
                       fc->valid && x_lp != nullptr && x_lp == fc->y_lp && fc->T == T && fc->B == B && fc->cols == D &&
                       (D % 8) == 0 && fc->nchunks > 1;
  if (chained) {
    cudaStream_t ss = fc->s;
    B2_CUDA(cudaStreamWaitEvent(ss, fc->ev_launch, 0));      // counters zeroed, weights of this step final
    rc = pack_weights(d, fw, bw, w, true, ss);
    if (rc) return rc;
    // completion order: chunk k is complete when fw has passed k+1 chunks and bw nchunks-k
    int order[kMaxChunks];
    const int K = fc->nchunks;
    for (int k = 0; k < K; ++k) order[k] = k;
    for (int i = 1; i < K; ++i)
      for (int j = i; j > 0; --j) {
        const int a0 = order[j - 1], a1 = order[j];
        const int c0 = (a0 + 1 > K - a0) ? a0 + 1 : K - a0, c1 = (a1 + 1 > K - a1) ? a1 + 1 : K - a1;
        if (c1 < c0) { order[j - 1] = a1; order[j] = a0; } else break;
      }
    const unsigned* prog = fc->progress[fc->pp];
    for (int i = 0; i < K && !rc; ++i) {
      const int k = order[i];
      const int t0 = k * fc->chunk_T, t1 = (t0 + fc->chunk_T < T) ? t0 + fc->chunk_T : T;
      const size_t r0 = (size_t)t0 * B;
      if (stream_wait_value_fn()(ss, (unsigned long long)(uintptr_t)(prog + k), fc->expected, 0u /* CU_STREAM_WAIT_VALUE_GEQ */) != 0) {
        set_error("tc_layer_forward: cuStreamWaitValue32 failed");
        rc = B2_ERR_CUDA;
        break;
      }
      // the two outermost chunks complete only when the recurrence kernel ends: no SM cap for them
      gemm_set_cta_limit(i + 2 < K ? fc->free_sms : 0);
      rc = gemm_bf16_tc(0, 1, (t1 - t0) * B, 8 * H, D, 1.f, x_lp + r0 * D, D, w.wx, 8 * H, Gbuf + r0 * 8 * H, 8 * H,
                        w.bias, EPI_STORE_F32, 0, ss);
    }
    gemm_set_cta_limit(0);
    if (rc) return rc;
    B2_CUDA(cudaEventRecord(fc->ev_done, ss));
    B2_CUDA(cudaStreamWaitEvent(stream, fc->ev_done, 0));
  } else {
    rc = pack_weights(d, fw, bw, w, r.wpack != nullptr, stream);
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
    rc = gemm_bf16_tc(0, 1, TB, 8 * H, D, 1.f, xa, ldx, w.wx, 8 * H, Gbuf, 8 * H, w.bias,
                      EPI_STORE_F32, 0, stream);
    if (rc) return rc;
  }
  fc->valid = false;
  RecFwdArgs ra;
  ra.progress = nullptr; ra.chunk_T = 1;
  if (want_chunks > 1 && !dbg_run && r.y_lp && T >= 4 * want_chunks) {
    const int K0 = want_chunks < kMaxChunks ? want_chunks : kMaxChunks;
    const int Tc = cdiv(T, K0);
    fc->pp ^= 1;
    B2_CUDA(cudaMemsetAsync(fc->progress[fc->pp], 0, kMaxChunks * sizeof(unsigned), stream));
    B2_CUDA(cudaEventRecord(fc->ev_launch, stream));
    ra.progress = fc->progress[fc->pp]; ra.chunk_T = Tc;
    const int ng = cdiv(B, RN);
    const int nch_env = env_int("B2_REC_NCHAIN", 0);
    const int nch = nch_env > 0 ? (nch_env > 2 ? 2 : nch_env) : (ng >= 2 ? 2 : 1);
    const int rec_ctas = 2 * cdiv(ng, nch) * (H / RU);
    fc->valid = true; fc->y_lp = r.y_lp; fc->T = T; fc->B = B; fc->cols = 2 * H; fc->chunk_T = Tc;
    fc->nchunks = cdiv(T, Tc); fc->expected = (unsigned)(2 * ng * (H / RU));
    fc->free_sms = num_sms() - rec_ctas > 8 ? num_sms() - rec_ctas : 8;
  }
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
    rc = rec_tc_forward(ra, Gbuf, env_int("B2_REC_NCHAIN", 0), env_int("B2_REC_GW", 0), stream);
    long long hb[48];
    cudaMemcpyAsync(hb, dbg_buf, sizeof(hb), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    fprintf(stderr, "[rec fwd dbg3] loop cycles per (cluster,chain): %lld %lld | %lld %lld | %lld %lld | %lld %lld\n",
            hb[16], hb[17], hb[18], hb[19], hb[20], hb[21], hb[22], hb[23]);
    fprintf(stderr, "[rec fwd dbg4] the same loops in ns (globaltimer): %lld %lld | %lld %lld -> clock64 rate %.3f GHz\n",
            hb[32], hb[33], hb[34], hb[35], hb[32] > 0 ? (double)hb[16] / (double)hb[32] : 0.0);
    fprintf(stderr, "[rec fwd dbg] cycles/step (needs a -DB2_REC_TIMING=1 build): reserve_stores=%lld | wait_acc=%lld "
            "ld+transpose=%lld wait_G=%lld math+stage=%lld fence+bar=%lld arrive+send=%lld\n",
            hb[1] / T, hb[2] / T, hb[3] / T, hb[4] / T, hb[5] / T, hb[6] / T, hb[7] / T);
    return rc;
  }
  prof_record(0, stream);
  rc = rec_tc_forward(ra, Gbuf, env_int("B2_REC_NCHAIN", 0), env_int("B2_REC_GW", 0), stream);
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
  // side-stream overlap needs the per-layer scratch in the reserve (a forward call made with need_backward)
  const bool use_side = env_int("B2_SIDE_STREAM", 1) != 0 && stream_wait_value_fn() != nullptr && r.dG != nullptr;
  SideCtx* sc = side_ctx();
  if (!use_side && (sc->def.valid || sc->pending)) { int rcj = tc_backward_join(stream); if (rcj) return rcj; }
  __nv_bfloat16* dG = r.dG ? r.dG : w.dG;
  float* dwx = r.dwx ? r.dwx : w.dwx; float* dwh = r.dwh ? r.dwh : w.dwh; float* dbias = r.dbias ? r.dbias : w.dbias;
  int rc = B2_OK;
  if (r.wpack) use_reserve_pack(d, r.wpack, &w);      // packed by this layer's forward call
  else rc = pack_weights(d, fw, bw, w, true, stream);
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
  if (d->dy_premasked) {
    ba.keep_prob = 1.f;          // the layer above applied this layer's mask in its dX GEMM epilogue
  } else if (d->keep_prob < 1.f && (((size_t)TB * 2 * H) % 4 == 0) && env_int("B2_DY_MASK_PASS", 1)) {     // mask dy once, outside the recurrence
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
  ba.wait_mode = env_int("B2_REC_WAIT", 0);
  ba.dbg = nullptr;
  const int nchain = env_int("B2_REC_NCHAIN", 0);
  const int ng = cdiv(B, RN);
  const int nch = nchain > 0 ? (nchain > 2 ? 2 : nchain) : (ng >= 2 ? 2 : 1);
  const int nclusters = 2 * cdiv(ng, nch);
  ba.resident = use_side ? sc->resident : nullptr;
  // dX chunk by chunk beside the BPTT kernel (same scheme as the forward chunk chain, inside one call)
  FwdChain* fc = fwd_chain();
  fc->valid = false;
  const int want_chunks = env_int("B2_BWD_CHUNKS", 16);
  const bool chunk_dx = want_chunks > 1 && dx != nullptr && stream_wait_value_fn() != nullptr && T >= 4 * want_chunks &&
                        !env_int("B2_REC_DBG", 0);
  int dx_chunk_T = T, dx_nchunks = 1;
  ba.progress = nullptr; ba.chunk_T = 1;
  if (chunk_dx) {
    dx_chunk_T = cdiv(T, want_chunks < kMaxChunks ? want_chunks : kMaxChunks);
    dx_nchunks = cdiv(T, dx_chunk_T);
    fc->pp ^= 1;
    B2_CUDA(cudaMemsetAsync(fc->progress[fc->pp], 0, kMaxChunks * sizeof(unsigned), stream));
    B2_CUDA(cudaEventRecord(fc->ev_launch, stream));
    ba.progress = fc->progress[fc->pp]; ba.chunk_T = dx_chunk_T;
  }
  if (env_int("B2_REC_DBG", 0)) {          // phase timers: needs the B2_BUILD_VARIANT=timing library
    static long long* dbg_buf = nullptr;
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 64 * sizeof(long long), stream);
    ba.dbg = dbg_buf;
    rc = rec_tc_backward(ba, dy, nchain, env_int("B2_REC_GW", 0), stream);
    long long hb[12];
    cudaMemcpyAsync(hb, dbg_buf, sizeof(hb), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    fprintf(stderr, "[rec bwd dbg] cycles/step: wait_ring=%lld coeffs=%lld wait_partials+sum=%lld dh_math=%lld stage=%lld "
            "bar1+dG=%lld wait_mma=%lld ld+convert+stage=%lld send=%lld\n",
            hb[1] / T, hb[8] / T, hb[0] / T, hb[9] / T, hb[10] / T, hb[3] / T, hb[4] / T, hb[5] / T, hb[6] / T);
  } else {
    prof_record(2, stream);
    rc = rec_tc_backward(ba, dy, nchain, env_int("B2_REC_GW", 0), stream);
    prof_record(3, stream);
  }
  if (rc) return rc;
  if (use_side) {
    sc->expected += (unsigned)nclusters;
    if (sc->def.valid) {
      // the previous layer's weight gradients: start once every cluster of this recurrence is resident,
      // on the SMs the clusters leave free
      const int r = stream_wait_value_fn()(sc->s, (unsigned long long)(uintptr_t)sc->resident, sc->expected, 0u /* GEQ */);
      if (r != 0) { set_error("cuStreamWaitValue32 failed (%d)", r); return B2_ERR_CUDA; }
      int free_sms = num_sms() - nclusters * (H / RU);
      if (free_sms < 16) free_sms = 16;
      rc = run_wgrad(sc->def, sc->s, free_sms);
      if (rc) return rc;
      B2_CUDA(cudaEventRecord(sc->ev_done, sc->s));
      sc->pending = true;
      sc->def.valid = false;
    }
    B2_CUDA(cudaEventRecord(sc->ev_rec, stream));          // dG of this layer complete (for a join-time flush)
  }
  // 2. critical path: dX[TB, D] = dG[TB, 8H] . Wx_packed^T  (sums both directions)
  const __nv_bfloat16* xa = x_lp;
  int ldx = D;
  if (!xa || (D % 8)) {
    ldx = (int)pad8(D);
    rc = cast_f32_bf16(x, TB, D, D, w.xb, ldx, stream);
    if (rc) return rc;
    xa = w.xb;
  }
  // DropoutWrapper backward of the layer below, fused into the store of dX (its own call then skips the mask pass)
  const bool dx_mask = dx != nullptr && d->dx_keep_prob > 0.f && d->dx_keep_prob < 1.f;
  if (dx && chunk_dx && dx_nchunks > 1) {
    cudaStream_t ss = fc->s;
    B2_CUDA(cudaStreamWaitEvent(ss, fc->ev_launch, 0));      // counters zeroed, dx buffer free
    const int K = dx_nchunks;
    int order[kMaxChunks];
    for (int k = 0; k < K; ++k) order[k] = k;
    for (int i = 1; i < K; ++i)
      for (int j = i; j > 0; --j) {
        const int a0 = order[j - 1], a1 = order[j];
        const int c0 = (a0 + 1 > K - a0) ? a0 + 1 : K - a0, c1 = (a1 + 1 > K - a1) ? a1 + 1 : K - a1;
        if (c1 < c0) { order[j - 1] = a1; order[j] = a0; } else break;
      }
    const unsigned expected = (unsigned)(2 * ng * (H / RU));
    int free_sms = num_sms() - nclusters * (H / RU);
    if (free_sms < 16) free_sms = 16;
    for (int i = 0; i < K && !rc; ++i) {
      const int k = order[i];
      const int t0 = k * dx_chunk_T, t1 = (t0 + dx_chunk_T < T) ? t0 + dx_chunk_T : T;
      const size_t r0 = (size_t)t0 * B;
      if (stream_wait_value_fn()(ss, (unsigned long long)(uintptr_t)(ba.progress + k), expected, 0u /* CU_STREAM_WAIT_VALUE_GEQ */) != 0) {
        set_error("tc_layer_backward: cuStreamWaitValue32 failed");
        rc = B2_ERR_CUDA;
        break;
      }
      gemm_set_cta_limit(i + 2 < K ? free_sms : 0);
      if (dx_mask) gemm_set_store_dropout(d->dx_keep_prob, d->dx_dropout_seed, (long long)r0);
      rc = gemm_bf16_tc(0, 0, (t1 - t0) * B, D, 8 * H, 1.f, dG + r0 * 8 * H, 8 * H, w.wx, 8 * H, dx + r0 * D, D, nullptr,
                        EPI_STORE_F32, 0, ss);
    }
    gemm_set_cta_limit(0);
    gemm_set_store_dropout(1.f, 0, 0);
    if (rc) return rc;
    B2_CUDA(cudaEventRecord(fc->ev_done, ss));
    B2_CUDA(cudaStreamWaitEvent(stream, fc->ev_done, 0));
  } else if (dx) {
    if (dx_mask) gemm_set_store_dropout(d->dx_keep_prob, d->dx_dropout_seed, 0);
    rc = gemm_bf16_tc(0, 0, TB, D, 8 * H, 1.f, dG, 8 * H, w.wx, 8 * H, dx, D, nullptr,
                      EPI_STORE_F32, 0, stream);
    gemm_set_store_dropout(1.f, 0, 0);
    if (rc) return rc;
  }
  // 3. off the critical path: weight gradients on packed operands
  DeferredWgrad wg;
  wg.valid = true; wg.T = T; wg.B = B; wg.D = D; wg.H = H; wg.ldx = ldx;
  wg.xa = xa; wg.dG = dG; wg.hs_lp = r.hs_lp; wg.dwx = dwx; wg.dwh = dwh; wg.dbias = dbias;
  wg.gk0 = g_fw->kernel; wg.gk1 = g_bw->kernel; wg.gb0 = g_fw->bias; wg.gb1 = g_bw->bias;
  // (a bf16 copy of x made in the shared workspace does not survive until a deferred launch)
  if (!use_side || xa == w.xb) {
    rc = run_wgrad(wg, stream, 0);
    if (rc) return rc;
    return dx ? B2_OK : tc_backward_join(stream);
  }
  sc->def = wg;
  // the first layer is the last one of a backward pass: hand the gradients back
  if (!dx) return tc_backward_join(stream);
  return B2_OK;
}

}  // namespace b2
