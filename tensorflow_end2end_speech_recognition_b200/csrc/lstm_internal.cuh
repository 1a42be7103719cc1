// Shared between lstm.cu (fp32 path + API dispatch), lstm_tc.cu (bf16 orchestration) and
// lstm_rec_tc.cu (tcgen05 recurrence kernels).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace b2 {

// reserve (saved by forward for backward), one per layer:
//   gates [T][B][2][H][4] f32  post-activation i, g, f, o of every cell (gate innermost)
//   cs    [T][B][2][H]    f32  cell state after step t (carried through inactive steps)
//   hs    [T][B][2][H]    f32  emitted h before dropout, fp32 path only
//   hs_lp [T*B][2H]       bf16 same, bf16 path (operand of the dWh GEMM)
//   y_lp  [T*B][2H]       bf16 layer output after dropout (= next layer's GEMM operand);
//                              aliases hs_lp when keep_prob == 1
bool tc_layer_supported(const b2_lstm_desc* d);

struct Reserve {
  float* gates; float* cs; float* hs;
  __nv_bfloat16* hs_lp; __nv_bfloat16* y_lp;
  float* hps;      // [T][B][2][P] projected h before dropout (num_proj > 0; `hs` then holds o*tanh(c))
  void* wpack;     // bf16 path: the layer's packed weights (wx | bias | wh | whT), written by forward and
                   // read back by backward, so the pack kernel runs once per layer and step
  // bf16 path, backward scratch that must outlive the layer's own call: the weight-gradient GEMMs of layer l
  // run on a side stream next to layer l-1's BPTT, so their operands cannot sit in the workspace all layers share
  __nv_bfloat16* dG;   // [T*B, 8H] gate gradients (packed order)
  float* dwx;          // [D, 8H]   packed input-weight gradient
  float* dwh;          // [2][H,4H] packed recurrent-weight gradient
  float* dbias;        // [8H]
};

// bytes of the packed-weight block kept in the reserve (bf16 path, need_backward)
inline size_t tc_pack_bytes(int D, int H) {
  size_t n = 0;
  n += align_up((size_t)D * 8 * H * 2, 1024);                              // wx   [D, 8H] bf16
  n += align_up((size_t)8 * H * 4, 1024);                                  // bias [8H] fp32
  n += align_up((size_t)2 * 4 * H * H * 2, 1024);                          // wh   [2][CS][128][H] bf16
  n += align_up((size_t)2 * (H / 32) * 4 * 128 * 128 * 2, 1024);           // whT  [2][CS][4][128][128] bf16
  return n;
}

bool wide_rec_supported(const b2_lstm_desc* d);
bool tc_layer_supported(const b2_lstm_desc* d);

inline size_t reserve_layout(const b2_lstm_desc* d, void* base, Reserve* r) {
  const size_t n = (size_t)d->T * d->B * 2 * d->H;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t og = take(n * 4 * sizeof(float));
  const size_t oc = take(n * sizeof(float));
  // bf16 shadows (h before dropout for the dWh GEMM, the emitted output for the next layer's GEMMs) instead of the fp32
  // hs when the tcgen05 recurrence or the grid-resident wide-layer recurrence runs: both kernels write them
  const bool lp = tc_layer_supported(d) || wide_rec_supported(d);
  const size_t oh = lp ? 0 : take(n * sizeof(float));
  const size_t ohl = lp ? take(n * 2) : 0;
  const size_t oyl = lp ? (d->keep_prob < 1.f ? take(n * 2) : ohl) : 0;
  const size_t ohp = d->num_proj > 0 ? take((size_t)d->T * d->B * 2 * d->num_proj * sizeof(float)) : 0;
  const bool lpb = tc_layer_supported(d) && d->need_backward;
  const size_t owp = lpb ? take(tc_pack_bytes(d->D_in, d->H)) : 0;
  const size_t odg = lpb ? take((size_t)d->T * d->B * 8 * d->H * 2) : 0;
  const size_t odwx = lpb ? take((size_t)d->D_in * 8 * d->H * 4) : 0;
  const size_t odwh = lpb ? take((size_t)2 * d->H * 4 * d->H * 4) : 0;
  const size_t odb = lpb ? take((size_t)8 * d->H * 4) : 0;
  if (r) {
    char* p = (char*)base;
    r->gates = (float*)(p + og); r->cs = (float*)(p + oc);
    r->hs = lp ? nullptr : (float*)(p + oh);
    r->hs_lp = lp ? (__nv_bfloat16*)(p + ohl) : nullptr;
    r->y_lp = lp ? (__nv_bfloat16*)(p + oyl) : nullptr;
    r->hps = d->num_proj > 0 ? (float*)(p + ohp) : nullptr;
    r->wpack = lpb ? (void*)(p + owp) : nullptr;
    r->dG = lpb ? (__nv_bfloat16*)(p + odg) : nullptr;
    r->dwx = lpb ? (float*)(p + odwx) : nullptr;
    r->dwh = lpb ? (float*)(p + odwh) : nullptr;
    r->dbias = lpb ? (float*)(p + odb) : nullptr;
  }
  return off;
}

int gemm_simt(int transa, int transb, int M, int N, int K, float alpha, const float* A, int lda,
              const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
              cudaStream_t stream);
int gemm_skinny_pair(int transb, int M, int N, int K, const float* A0, const float* A1, int lda,
                     const float* B0, const float* B1, int ldb, float* C0, float* C1, int ldc,
                     cudaStream_t stream);
int gemm_bf16_tc(int a_mn, int b_mn, int M, int N, int K, float alpha, const __nv_bfloat16* A,
                 int lda, const __nv_bfloat16* B, int ldb, void* C, int ldc, const float* bias,
                 int epi, int k_splits_hint, cudaStream_t stream);
int gemm_bf16_tc_conv(int M, int N, int Ktap, int taps, int tap_rows, const __nv_bfloat16* A, int64_t a_rows, int lda,
                      const __nv_bfloat16* B, int ldb, float* C, int ldc, cudaStream_t stream);
int cast_f32_bf16(const float* in, int64_t rows, int cols, int ldi, __nv_bfloat16* out, int ldo,
                  cudaStream_t stream);
int make_tmap_generic(CUtensorMap* tm, int dtype_is_f32, const void* base, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      int swizzle128);
int num_sms();
void gemm_set_cta_limit(int n);
void gemm_set_store_dropout(float keep, unsigned long long seed, long long row0);   // fused into the fp32 store of
                                                                                   // the next gemm_bf16_tc launches; keep = 1: off

// bf16 / tcgen05 layer (lstm_tc.cu)
size_t tc_layer_workspace_bytes(const b2_lstm_desc* d);
int tc_layer_forward(const b2_lstm_desc* d, const float* x, const __nv_bfloat16* x_lp,
                     const int32_t* seq_len, const b2_lstm_params* fw, const b2_lstm_params* bw,
                     float* y, float* final_state, void* reserve, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream);
int tc_layer_backward(const b2_lstm_desc* d, const float* x, const __nv_bfloat16* x_lp,
                      const int32_t* seq_len, const b2_lstm_params* fw, const b2_lstm_params* bw,
                      const float* dy, const float* d_final_state, const void* reserve, float* dx,
                      const b2_lstm_grads* g_fw,
                      const b2_lstm_grads* g_bw, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream);

// grid-resident recurrence for wide layers (lstm_wide.cu)
bool wide_rec_supported(const b2_lstm_desc* d);
size_t wide_rec_workspace_bytes(const b2_lstm_desc* d);
int wide_rec_forward(const b2_lstm_desc* d, const b2_lstm_params* fw, const b2_lstm_params* bw, const int32_t* seq_len,
                     const float* G, float* y, float* gates, float* cs, __nv_bfloat16* hs_lp, __nv_bfloat16* y_lp, float* final_state, void* workspace,
                     cudaStream_t stream);

int wide_rec_backward(const b2_lstm_desc* d, const b2_lstm_params* fw, const b2_lstm_params* bw, const int32_t* seq_len,
                      const float* dy, const float* gates, const float* cs, const float* d_final_state, float* dG,
                      __nv_bfloat16* dG_lp, void* workspace, cudaStream_t stream);
int tc_backward_join(cudaStream_t stream);
int tc_backward_side_wait(cudaStream_t stream);
void tc_profile_enable(int on);
int tc_profile_last_ms(float* fwd_ms, float* bwd_ms);

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace b2
