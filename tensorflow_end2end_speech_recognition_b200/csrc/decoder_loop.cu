// The attention decoder loop behind ONE C entry point per direction (forward: teacher-forced or
// greedy; backward: the teacher-forced loop differentiated).
//
// Replaces tf.while_loop driving AttentionDecoder.step (models/attention/decoders/
// dynamic_decoder.py:148-212 + attention_decoder.py:256-295): the reference's loop also runs in
// the framework's native runtime, not in Python.  One iteration = cell pre-activation GEMM ->
// gate math -> (query GEMM) -> attention step -> attentional-vector GEMMs + tanh -> logits GEMM ->
// arg-max -> emit (impute_finished, state copy-through, input feeding, embedding gather, finished
// flags).  ~12 launches per step issued back to back; no host synchronisation in the teacher-forced
// path, one poll of the finished flags every `poll_every` steps in the greedy path.
//
// Backward: everything that is not sequential is time-batched (output layer, attentional vector,
// cell-kernel / bias / peephole / embedding / W_query gradients, d(enc) through the context as one
// GEMM per utterance); the per-step remainder is attention backward -> query GEMM -> gate-math
// backward -> three cell-kernel GEMMs.  All decoder arithmetic is fp32 (CUDA-core GEMM: the
// matrices have B <= 64 rows).
#include "common.cuh"
#include "lstm_internal.cuh"

namespace b2 {

struct DecSaved {
  float* xh; float* z; float* c; float* h; float* alpha; float* ctx; float* av; float* q; float* energy;
};
static size_t dec_saved_layout(const b2_decoder_desc* d, int L, void* base, DecSaved* s) {
  const size_t B = d->B, X = (size_t)d->emb + d->E + d->Hd, Ls = L > 0 ? L : 1;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n * 4, 256); return o; };
  const size_t oxh = take((Ls + 1) * B * X), oz = take(Ls * B * 4 * d->Hd), oc = take((Ls + 1) * B * d->Hd);
  const size_t oh = take(Ls * B * d->Hd), oa = take(Ls * B * d->T), octx = take(Ls * B * d->E);
  const size_t oav = take(Ls * B * d->Hd);
  const size_t oq = d->query_projected ? take(Ls * B * d->A) : 0;
  const size_t oe = d->sigmoid_smoothing ? take(Ls * B * d->T) : 0;
  if (s) {
    char* p = (char*)base;
    s->xh = (float*)(p + oxh); s->z = (float*)(p + oz); s->c = (float*)(p + oc); s->h = (float*)(p + oh);
    s->alpha = (float*)(p + oa); s->ctx = (float*)(p + octx); s->av = (float*)(p + oav);
    s->q = d->query_projected ? (float*)(p + oq) : nullptr;
    s->energy = d->sigmoid_smoothing ? (float*)(p + oe) : nullptr;
  }
  return off;
}

// forward scratch when nothing is saved (inference): one step's worth of every buffer
struct DecScratch {
  float* xh; float* z; float* c_new; float* h_new; float* alpha; float* alpha2; float* ctx; float* av; float* q;
  float* logits; int* ids; float* zeros_alpha;
};
static size_t dec_scratch_layout(const b2_decoder_desc* d, void* base, DecScratch* s) {
  const size_t B = d->B, X = (size_t)d->emb + d->E + d->Hd;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n * 4, 256); return o; };
  const size_t a0 = take(2 * B * X), a1 = take(B * 4 * d->Hd), a2 = take(B * d->Hd), a3 = take(B * d->Hd);
  const size_t a4 = take(B * d->T), a4b = take(B * d->T), a5 = take(B * d->E), a6 = take(B * d->Hd);
  const size_t a7 = take(B * (d->A > d->Hd ? d->A : d->Hd));
  const size_t a8 = take(B * d->C), a9 = take(B), a10 = take(B * d->T);
  if (s) {
    char* p = (char*)base;
    s->xh = (float*)(p + a0); s->z = (float*)(p + a1); s->c_new = (float*)(p + a2); s->h_new = (float*)(p + a3);
    s->alpha = (float*)(p + a4); s->alpha2 = (float*)(p + a4b); s->ctx = (float*)(p + a5); s->av = (float*)(p + a6);
    s->q = (float*)(p + a7); s->logits = (float*)(p + a8); s->ids = (int*)(p + a9); s->zeros_alpha = (float*)(p + a10);
  }
  return off;
}

// xh0[b] = [embedding(first id) ; 0 ; h0[b]]
__global__ void __launch_bounds__(256)
decoder_init_kernel(const float* __restrict__ embedding, const int* __restrict__ labels, int labels_ld, int sos,
                    const float* __restrict__ h0, int emb, int E, int Hd, int C, float* __restrict__ xh) {
  const int b = blockIdx.x;
  const int X = emb + E + Hd;
  const int id = labels ? labels[(size_t)b * labels_ld] : sos;
  float* xr = xh + (size_t)b * X;
  for (int i = threadIdx.x; i < X; i += 256) {
    float v = 0.f;
    if (i < emb) v = (id >= 0 && id < C) ? embedding[(size_t)id * emb + i] : 0.f;
    else if (i >= emb + E) v = h0[(size_t)b * Hd + (i - emb - E)];
    xr[i] = v;
  }
}

__global__ void decoder_init_finished_kernel(const int* __restrict__ dec_len, int B, int max_iter, int* finished) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) finished[b] = dec_len ? (dec_len[b] <= 0) : (max_iter <= 0);
}

// time-major [L,B] token ids from batch-major labels [B, ld]
__global__ void decoder_ids_tm_kernel(const int* __restrict__ labels, int ld, int B, int L, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L * B) out[i] = labels[(size_t)(i % B) * ld + (i / B)];
}

}  // namespace b2

using namespace b2;

static int dec_check(const b2_decoder_desc* d) {
  B2_CHECK_ARG(d != nullptr, "decoder: null descriptor");
  B2_CHECK_ARG(d->B > 0 && d->T > 0 && d->E > 0 && d->Hd > 0 && d->A > 0 && d->emb > 0 && d->C > 0,
               "decoder: bad shape");
  B2_CHECK_ARG(d->E % 4 == 0, "decoder: encoder width must be a multiple of 4");
  B2_CHECK_ARG(d->query_projected || d->A == d->Hd, "decoder: unprojected query needs A == Hd");
  return B2_OK;
}

extern "C" size_t b2_attention_decoder_reserve_bytes(const b2_decoder_desc* d, int max_steps) {
  return d ? dec_saved_layout(d, max_steps, nullptr, nullptr) : 0;
}
extern "C" size_t b2_attention_decoder_workspace_bytes(const b2_decoder_desc* d, int max_steps) {
  if (!d) return 0;
  // forward scratch, or the backward's time-batched buffers (larger)
  const size_t L = max_steps > 0 ? max_steps : 1, B = d->B;
  size_t bwd = 0;
  auto add = [&](size_t n) { bwd += align_up(n * 4, 256); };
  add(L * B * d->Hd); add(L * B * d->Hd); add(L * B * d->E); add(L * B * d->A);
  add(L * B * 4 * d->Hd); add(L * B * d->emb); add(L * B); add(B * d->Hd);
  bwd += b2_attention_step_backward_workspace_bytes(d->B, d->T) + 256;
  const size_t fwd = dec_scratch_layout(d, nullptr, nullptr);
  return fwd > bwd ? fwd : bwd;
}

extern "C" int b2_attention_decoder_forward(const b2_decoder_desc* d, const b2_decoder_params* p,
                                            const float* enc, const float* keys, const int32_t* enc_len,
                                            const float* c0, const float* h0, const int32_t* labels,
                                            int labels_ld, const int32_t* dec_len, int sos, int eos,
                                            int max_steps, int poll_every, void* reserve,
                                            float* out_logits, int32_t* out_ids, float* out_av,
                                            float* out_alpha, float* out_ctx, float* c_state,
                                            float* h_state, int32_t* finished, int32_t* steps_run,
                                            void* workspace, size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = dec_check(d);
  if (rc) return rc;
  B2_CHECK_ARG(p && enc && enc_len && c0 && h0 && out_logits && out_ids && out_av && out_alpha && out_ctx &&
               c_state && h_state && finished && workspace, "b2_attention_decoder_forward: null pointer");
  B2_CHECK_ARG(!labels || (dec_len && labels_ld > 1), "b2_attention_decoder_forward: teacher forcing needs dec_len");
  B2_CHECK_ARG(!reserve || labels, "b2_attention_decoder_forward: saving for backward needs teacher forcing");
  B2_CHECK_ARG(!reserve || !d->feed_previous_attention,
               "b2_attention_decoder_forward: training with feed_previous_attention is not built");
  const int B = d->B, T = d->T, E = d->E, Hd = d->Hd, A = d->A, emb = d->emb, C = d->C;
  const int X = emb + E + Hd;
  const int L = max_steps;
  if (steps_run) *steps_run = 0;
  if (L <= 0) return B2_OK;
  DecScratch w;
  if (workspace_bytes < dec_scratch_layout(d, workspace, &w)) { set_error("b2_attention_decoder_forward: workspace too small"); return B2_ERR_WORKSPACE; }
  DecSaved sv;
  if (reserve) dec_saved_layout(d, L, reserve, &sv);
  const bool teacher = labels != nullptr;
  const bool loc = d->filter_width > 0;
  B2_CUDA(cudaMemcpyAsync(c_state, c0, (size_t)B * Hd * 4, cudaMemcpyDeviceToDevice, stream));
  B2_CUDA(cudaMemcpyAsync(h_state, h0, (size_t)B * Hd * 4, cudaMemcpyDeviceToDevice, stream));
  B2_CUDA(cudaMemsetAsync(w.zeros_alpha, 0, (size_t)B * T * 4, stream));
  float* xh = reserve ? sv.xh : w.xh;
  decoder_init_kernel<<<B, 256, 0, stream>>>(p->embedding, labels, labels_ld, sos, h0, emb, E, Hd, C, xh);
  B2_LAUNCH_CHECK();
  decoder_init_finished_kernel<<<cdiv(B, 128), 128, 0, stream>>>(dec_len, B, teacher ? 1 : L, finished);
  B2_LAUNCH_CHECK();
  if (reserve) B2_CUDA(cudaMemcpyAsync(sv.c, c0, (size_t)B * Hd * 4, cudaMemcpyDeviceToDevice, stream));
  const float* prev_alpha = nullptr;                 // NULL = all zero (b2_attention_step_forward)
  int* h_fin = nullptr;
  if (!teacher && poll_every > 0) B2_CUDA(cudaMallocHost(&h_fin, (size_t)B * sizeof(int)));
  int t = 0;
  for (; t < L; ++t) {
    float* z = reserve ? sv.z + (size_t)t * B * 4 * Hd : w.z;
    float* c_new = reserve ? sv.c + (size_t)(t + 1) * B * Hd : w.c_new;
    float* h_new = reserve ? sv.h + (size_t)t * B * Hd : w.h_new;
    float* alpha = reserve ? sv.alpha + (size_t)t * B * T : ((t & 1) ? w.alpha2 : w.alpha);
    float* ctx = reserve ? sv.ctx + (size_t)t * B * E : w.ctx;
    float* av = reserve ? sv.av + (size_t)t * B * Hd : w.av;
    float* q = d->query_projected ? (reserve ? sv.q + (size_t)t * B * A : w.q) : h_new;
    float* energy = (reserve && sv.energy) ? sv.energy + (size_t)t * B * T : nullptr;
    float* xh_next = reserve ? sv.xh + (size_t)(t + 1) * B * X : (w.xh + (size_t)((t + 1) & 1) * B * X);
    if ((rc = gemm_simt(0, 0, B, 4 * Hd, X, 1.f, xh, X, p->cell_kernel, 4 * Hd, 0.f, z, 4 * Hd, nullptr, stream))) break;
    if ((rc = b2_lstm_cell_pointwise(z, p->cell_bias, p->w_i_diag, p->w_f_diag, p->w_o_diag, c_state, B, Hd,
                                     d->forget_bias, d->cell_clip, c_new, h_new, stream_))) break;
    if (d->query_projected)
      if ((rc = gemm_simt(0, 0, B, A, Hd, 1.f, h_new, Hd, p->w_query, A, 0.f, q, A, nullptr, stream))) break;
    if ((rc = b2_attention_step_forward(d->attention_mode, enc, keys, q, prev_alpha, enc_len,
                                        loc ? p->conv_filter : nullptr, d->filter_width, p->w_filter, p->b_filter,
                                        p->v_a, B, T, E, A, d->sharpening, d->sigmoid_smoothing, alpha, ctx, energy,
                                        stream_))) break;
    if ((rc = gemm_simt(0, 0, B, Hd, Hd, 1.f, h_new, Hd, p->w_av, Hd, 0.f, av, Hd, nullptr, stream))) break;
    if ((rc = gemm_simt(0, 0, B, Hd, E, 1.f, ctx, E, p->w_av + (size_t)Hd * Hd, Hd, 1.f, av, Hd, nullptr, stream))) break;
    if ((rc = b2_tanh_inplace(av, (int64_t)B * Hd, stream_))) break;
    if ((rc = gemm_simt(0, 0, B, C, Hd, 1.f, av, Hd, p->w_out, C, 0.f, w.logits, C, p->b_out, stream))) break;
    if ((rc = b2_argmax_rows(w.logits, B, C, w.ids, stream_))) break;
    if ((rc = b2_decoder_step_emit(B, C, Hd, E, T, emb, t, L, w.logits, w.ids, av, alpha, ctx, c_new, h_new,
                                   c_state, h_state, finished, p->embedding, labels, labels_ld, dec_len,
                                   teacher ? -1 : eos, teacher ? 0 : L, xh_next, out_logits, out_ids, out_av,
                                   out_alpha, out_ctx, stream_))) break;
    xh = xh_next;
    if (d->feed_previous_attention) prev_alpha = alpha;
    if (h_fin && (t + 1) % poll_every == 0 && t + 1 < L) {
      cudaMemcpyAsync(h_fin, finished, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, stream);
      cudaStreamSynchronize(stream);
      bool all = true;
      for (int b = 0; b < B; ++b) all = all && h_fin[b];
      if (all) { ++t; break; }
    }
  }
  if (h_fin) cudaFreeHost(h_fin);
  if (steps_run) *steps_run = t < L ? t : L;
  return rc;
}

extern "C" int b2_attention_decoder_backward(const b2_decoder_desc* d, const b2_decoder_params* p,
                                             const float* enc, const float* keys, const int32_t* enc_len,
                                             const int32_t* labels, int labels_ld, int steps,
                                             const void* reserve, const float* dlogits_tm,
                                             const b2_decoder_grads* g, float* d_keys, float* d_enc,
                                             float* dc0, float* dh0, void* workspace,
                                             size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = dec_check(d);
  if (rc) return rc;
  B2_CHECK_ARG(p && enc && enc_len && labels && reserve && dlogits_tm && g && d_enc && dc0 && dh0 && workspace,
               "b2_attention_decoder_backward: null pointer");
  B2_CHECK_ARG(steps > 0, "b2_attention_decoder_backward: no steps");
  B2_CHECK_ARG(!d->feed_previous_attention, "b2_attention_decoder_backward: feed_previous_attention not built");
  if (workspace_bytes < b2_attention_decoder_workspace_bytes(d, steps)) { set_error("b2_attention_decoder_backward: workspace too small"); return B2_ERR_WORKSPACE; }
  const int B = d->B, T = d->T, E = d->E, Hd = d->Hd, A = d->A, emb = d->emb, C = d->C, L = steps;
  const int X = emb + E + Hd;
  const int64_t LB = (int64_t)L * B;
  DecSaved sv;
  dec_saved_layout(d, L, (void*)reserve, &sv);
  // workspace carve-up (same order as b2_attention_decoder_workspace_bytes)
  char* wp = (char*)workspace;
  auto take = [&](size_t n) { float* r = (float*)wp; wp += align_up(n * 4, 256); return r; };
  float* d_av = take((size_t)LB * Hd);
  float* dh_av = take((size_t)LB * Hd);
  float* dctx_all = take((size_t)LB * E);
  float* dq_all = take((size_t)LB * A);
  float* dz_all = take((size_t)LB * 4 * Hd);
  float* demb_all = take((size_t)LB * emb);
  int* ids_tm = (int*)take((size_t)LB);
  float* dc_buf = take((size_t)B * Hd);
  void* att_ws = (void*)wp;
  const size_t att_ws_bytes = b2_attention_step_backward_workspace_bytes(B, T);
  const float* dl = dlogits_tm;
  // ---- time-batched head
  if ((rc = gemm_simt(1, 0, Hd, C, (int)LB, 1.f, sv.av, Hd, dl, C, 1.f, g->w_out, C, nullptr, stream))) return rc;
  if ((rc = b2_colsum(dl, LB, C, C, g->b_out, 1, stream_))) return rc;
  if ((rc = gemm_simt(0, 1, (int)LB, Hd, C, 1.f, dl, C, p->w_out, C, 0.f, d_av, Hd, nullptr, stream))) return rc;
  if ((rc = b2_tanh_backward(d_av, sv.av, d_av, LB * Hd, stream_))) return rc;                 // d_pre in place
  if ((rc = gemm_simt(1, 0, Hd, Hd, (int)LB, 1.f, sv.h, Hd, d_av, Hd, 1.f, g->w_av, Hd, nullptr, stream))) return rc;
  if ((rc = gemm_simt(1, 0, E, Hd, (int)LB, 1.f, sv.ctx, E, d_av, Hd, 1.f, g->w_av + (size_t)Hd * Hd, Hd, nullptr, stream))) return rc;
  if ((rc = gemm_simt(0, 1, (int)LB, Hd, Hd, 1.f, d_av, Hd, p->w_av, Hd, 0.f, dh_av, Hd, nullptr, stream))) return rc;
  if ((rc = gemm_simt(0, 1, (int)LB, E, Hd, 1.f, d_av, Hd, p->w_av + (size_t)Hd * Hd, Hd, 0.f, dctx_all, E, nullptr, stream))) return rc;
  // ---- sequential part
  const float* k_emb = p->cell_kernel;
  const float* k_ctx = p->cell_kernel + (size_t)emb * 4 * Hd;
  const float* k_h = p->cell_kernel + (size_t)(emb + E) * 4 * Hd;
  const bool loc = d->filter_width > 0;
  const float* dc_in = nullptr;
  float* dc_a = dc_buf; float* dc_b = dc0;          // ping-pong so the last write lands in dc0 or is copied
  for (int t = L - 1; t >= 0; --t) {
    float* dh_t = dh_av + (size_t)t * B * Hd;
    float* dctx_t = dctx_all + (size_t)t * B * E;
    const float* alpha = sv.alpha + (size_t)t * B * T;
    const float* energy = sv.energy ? sv.energy + (size_t)t * B * T : nullptr;
    if (d->query_projected) {
      float* dq = dq_all + (size_t)t * B * A;
      if ((rc = b2_attention_step_backward(d->attention_mode, enc, keys, sv.q + (size_t)t * B * A, alpha, energy, enc_len,
                                           loc ? p->b_filter : nullptr, p->v_a, B, T, E, A, d->sharpening,
                                           d->sigmoid_smoothing, dctx_t, d_keys, dq, 0, g->v_a,
                                           loc ? g->b_filter : nullptr, att_ws, att_ws_bytes, stream_))) return rc;
      if ((rc = gemm_simt(0, 1, B, Hd, A, 1.f, dq, A, p->w_query, A, 1.f, dh_t, Hd, nullptr, stream))) return rc;
    } else {
      if ((rc = b2_attention_step_backward(d->attention_mode, enc, keys, sv.h + (size_t)t * B * Hd, alpha, energy, enc_len,
                                           loc ? p->b_filter : nullptr, p->v_a, B, T, E, A, d->sharpening,
                                           d->sigmoid_smoothing, dctx_t, d_keys, dh_t, 1, g->v_a,
                                           loc ? g->b_filter : nullptr, att_ws, att_ws_bytes, stream_))) return rc;
    }
    float* dz = dz_all + (size_t)t * B * 4 * Hd;
    float* dc_out = (t & 1) ? dc_a : dc_b;
    if ((rc = b2_lstm_cell_pointwise_backward(sv.z + (size_t)t * B * 4 * Hd, p->cell_bias, p->w_i_diag, p->w_f_diag,
                                              p->w_o_diag, sv.c + (size_t)t * B * Hd, dh_t, dc_in, B, Hd,
                                              d->forget_bias, d->cell_clip, dz, dc_out, stream_))) return rc;
    dc_in = dc_out;
    if ((rc = gemm_simt(0, 1, B, emb, 4 * Hd, 1.f, dz, 4 * Hd, k_emb, 4 * Hd, 0.f, demb_all + (size_t)t * B * emb, emb, nullptr, stream))) return rc;
    if (t > 0) {
      if ((rc = gemm_simt(0, 1, B, E, 4 * Hd, 1.f, dz, 4 * Hd, k_ctx, 4 * Hd, 1.f, dctx_all + (size_t)(t - 1) * B * E, E, nullptr, stream))) return rc;
      if ((rc = gemm_simt(0, 1, B, Hd, 4 * Hd, 1.f, dz, 4 * Hd, k_h, 4 * Hd, 1.f, dh_av + (size_t)(t - 1) * B * Hd, Hd, nullptr, stream))) return rc;
    } else {
      if ((rc = gemm_simt(0, 1, B, Hd, 4 * Hd, 1.f, dz, 4 * Hd, k_h, 4 * Hd, 0.f, dh0, Hd, nullptr, stream))) return rc;
    }
  }
  if (dc_in != dc0) B2_CUDA(cudaMemcpyAsync(dc0, dc_in, (size_t)B * Hd * 4, cudaMemcpyDeviceToDevice, stream));
  // ---- time-batched tails
  decoder_ids_tm_kernel<<<cdiv(LB, 256), 256, 0, stream>>>(labels, labels_ld, B, L, ids_tm);
  B2_LAUNCH_CHECK();
  if ((rc = b2_embedding_grad(demb_all, emb, ids_tm, LB, emb, C, g->embedding, stream_))) return rc;
  if ((rc = gemm_simt(1, 0, X, 4 * Hd, (int)LB, 1.f, sv.xh, X, dz_all, 4 * Hd, 1.f, g->cell_kernel, 4 * Hd, nullptr, stream))) return rc;
  if ((rc = b2_colsum(dz_all, LB, 4 * Hd, 4 * Hd, g->cell_bias, 1, stream_))) return rc;
  if (p->w_i_diag)
    if ((rc = b2_decoder_peephole_grad(dz_all, sv.c, L, B, Hd, g->w_i_diag, g->w_f_diag, g->w_o_diag, stream_))) return rc;
  if (d->query_projected)
    if ((rc = gemm_simt(1, 0, Hd, A, (int)LB, 1.f, sv.h, Hd, dq_all, A, 1.f, g->w_query, A, nullptr, stream))) return rc;
  // d(enc) through the context: d_enc[b] += Alpha[b]^T . Dctx[b]   ([T x L] . [L x E])
  for (int b = 0; b < B; ++b)
    if ((rc = gemm_simt(1, 0, T, E, L, 1.f, sv.alpha + (size_t)b * T, B * T, dctx_all + (size_t)b * E, B * E, 1.f,
                        d_enc + (size_t)b * T * E, E, nullptr, stream))) return rc;
  return B2_OK;
}
