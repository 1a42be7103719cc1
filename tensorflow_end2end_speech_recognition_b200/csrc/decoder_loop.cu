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

// pinned host words for the finished-flag polls of the greedy / beam loops (grown on demand, per thread)
static int* poll_buffer(int n) {
  static thread_local int* buf = nullptr;
  static thread_local int cap = 0;
  if (n > cap) {
    if (buf) cudaFreeHost(buf);
    buf = nullptr; cap = 0;
    if (cudaMallocHost(&buf, (size_t)n * sizeof(int)) != cudaSuccess) { buf = nullptr; return nullptr; }
    cap = n;
  }
  return buf;
}

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
  B2_CHECK_ARG(d->keep_prob_decoder > 0.f && d->keep_prob_decoder <= 1.f && d->keep_prob_embedding > 0.f &&
               d->keep_prob_embedding <= 1.f, "decoder: keep_prob out of (0,1]");
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
  add(L * B * 4 * d->Hd); add(L * B * d->emb); add(L * B); add(B * d->Hd); add(B * d->Hd); add(2 * B * d->T);
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
  // dropout only in the training pass (teacher forcing with a reserve)
  const float kd = reserve ? d->keep_prob_decoder : 1.f, ke = reserve ? d->keep_prob_embedding : 1.f;
  if (ke < 1.f)
    if ((rc = b2_dropout_rows(xh, X, xh, X, B, emb, ke, d->dropout_seed + 1, 0, (uint64_t)labels_ld * emb, 0, stream_))) return rc;
  const float* prev_alpha = nullptr;                 // NULL = all zero (b2_attention_step_forward)
  int* h_fin = (!teacher && poll_every > 0) ? poll_buffer(B) : nullptr;
  int t = 0;
  for (; t < L; ++t) {
    float* z = reserve ? sv.z + (size_t)t * B * 4 * Hd : w.z;
    float* c_new = reserve ? sv.c + (size_t)(t + 1) * B * Hd : w.c_new;
    float* h_new = (reserve && kd >= 1.f) ? sv.h + (size_t)t * B * Hd : w.h_new;   // cell state h (never dropped)
    float* alpha = reserve ? sv.alpha + (size_t)t * B * T : ((t & 1) ? w.alpha2 : w.alpha);
    float* ctx = reserve ? sv.ctx + (size_t)t * B * E : w.ctx;
    float* av = reserve ? sv.av + (size_t)t * B * Hd : w.av;
    // cell OUTPUT = DropoutWrapper(h): what the attention and the attentional vector see
    float* h_use = h_new;
    if (kd < 1.f) h_use = sv.h + (size_t)t * B * Hd;
    float* q = d->query_projected ? (reserve ? sv.q + (size_t)t * B * A : w.q) : h_use;
    float* energy = (reserve && sv.energy) ? sv.energy + (size_t)t * B * T : nullptr;
    float* xh_next = reserve ? sv.xh + (size_t)(t + 1) * B * X : (w.xh + (size_t)((t + 1) & 1) * B * X);
    if ((rc = gemm_simt(0, 0, B, 4 * Hd, X, 1.f, xh, X, p->cell_kernel, 4 * Hd, 0.f, z, 4 * Hd, nullptr, stream))) break;
    if ((rc = b2_lstm_cell_pointwise(z, p->cell_bias, p->w_i_diag, p->w_f_diag, p->w_o_diag, c_state, B, Hd,
                                     d->forget_bias, d->cell_clip, c_new, h_new, stream_))) break;
    if (kd < 1.f)
      if ((rc = b2_dropout_rows(h_new, Hd, h_use, Hd, B, Hd, kd, d->dropout_seed, (uint64_t)t * B * Hd, Hd, 0, stream_))) break;
    if (d->query_projected)
      if ((rc = gemm_simt(0, 0, B, A, Hd, 1.f, h_use, Hd, p->w_query, A, 0.f, q, A, nullptr, stream))) break;
    if ((rc = b2_attention_step_forward(d->attention_mode, enc, keys, q, prev_alpha, enc_len,
                                        loc ? p->conv_filter : nullptr, d->filter_width, p->w_filter, p->b_filter,
                                        p->v_a, B, T, E, A, d->sharpening, d->sigmoid_smoothing, alpha, ctx, energy,
                                        stream_))) break;
    if ((rc = gemm_simt(0, 0, B, Hd, Hd, 1.f, h_use, Hd, p->w_av, Hd, 0.f, av, Hd, nullptr, stream))) break;
    if ((rc = gemm_simt(0, 0, B, Hd, E, 1.f, ctx, E, p->w_av + (size_t)Hd * Hd, Hd, 1.f, av, Hd, nullptr, stream))) break;
    if ((rc = b2_tanh_inplace(av, (int64_t)B * Hd, stream_))) break;
    if ((rc = gemm_simt(0, 0, B, C, Hd, 1.f, av, Hd, p->w_out, C, 0.f, w.logits, C, p->b_out, stream))) break;
    if ((rc = b2_argmax_rows(w.logits, B, C, w.ids, stream_))) break;
    if ((rc = b2_decoder_step_emit(B, C, Hd, E, T, emb, t, L, w.logits, w.ids, av, alpha, ctx, c_new, h_new,
                                   c_state, h_state, finished, p->embedding, labels, labels_ld, dec_len,
                                   teacher ? -1 : eos, teacher ? 0 : L, xh_next, out_logits, out_ids, out_av,
                                   out_alpha, out_ctx, stream_))) break;
    if (ke < 1.f && t + 1 < L)        // the next input's embedding is labels_embedded[:, t+1] with ITS dropout mask
      if ((rc = b2_dropout_rows(xh_next, X, xh_next, X, B, emb, ke, d->dropout_seed + 1, (uint64_t)(t + 1) * emb,
                                (uint64_t)labels_ld * emb, 0, stream_))) break;
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
  float* dc_tmp = take((size_t)B * Hd);
  float* dal = take((size_t)2 * B * T);          // d(alpha) exchanged between steps (location term, real previous weights)
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
  const float kd = d->keep_prob_decoder, ke = d->keep_prob_embedding;
  // DropoutWrapper on the cell output: sv.h holds the DROPPED output (operand of the query / attentional-vector
  // products); its gradient passes the same mask.  Masking is linear, so the attentional-vector share is masked
  // here, the attention share per step below; the recurrent share (cell state h, never dropped) is not masked.
  if (kd < 1.f)
    if ((rc = b2_dropout_rows(dh_av, Hd, dh_av, Hd, LB, Hd, kd, d->dropout_seed, 0, Hd, 0, stream_))) return rc;
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
    // location term fed by the previous step's weights: its gradient reaches alpha_{t-1} (handed to step t-1
    // through dal) as well as the conv filter and W_filter.  Step 0 saw zeros (no conv, nothing to hand on).
    const bool fp = loc && d->feed_previous_attention;
    const float* prev_al = (fp && t > 0) ? sv.alpha + (size_t)(t - 1) * B * T : nullptr;
    const float* dal_in = (fp && t < L - 1) ? dal + (size_t)((t + 1) & 1) * B * T : nullptr;
    float* dal_out = prev_al ? dal + (size_t)(t & 1) * B * T : nullptr;
    if (d->query_projected) {
      float* dq = dq_all + (size_t)t * B * A;
      if ((rc = b2_attention_step_backward_loc(d->attention_mode, enc, keys, sv.q + (size_t)t * B * A, alpha, energy, enc_len,
                                               loc ? p->b_filter : nullptr, p->v_a, B, T, E, A, d->sharpening,
                                               d->sigmoid_smoothing, dctx_t, d_keys, dq, 0, g->v_a,
                                               loc ? g->b_filter : nullptr, prev_al, p->conv_filter, d->filter_width,
                                               p->w_filter, dal_in, dal_out, g->conv_filter, g->w_filter,
                                               att_ws, att_ws_bytes, stream_))) return rc;
      if (kd < 1.f) {
        if ((rc = gemm_simt(0, 1, B, Hd, A, 1.f, dq, A, p->w_query, A, 0.f, dc_tmp, Hd, nullptr, stream))) return rc;
        if ((rc = b2_dropout_rows(dc_tmp, Hd, dh_t, Hd, B, Hd, kd, d->dropout_seed, (uint64_t)t * B * Hd, Hd, 1, stream_))) return rc;
      } else {
        if ((rc = gemm_simt(0, 1, B, Hd, A, 1.f, dq, A, p->w_query, A, 1.f, dh_t, Hd, nullptr, stream))) return rc;
      }
    } else {
      float* dq_dst = kd < 1.f ? dc_tmp : dh_t;
      if ((rc = b2_attention_step_backward_loc(d->attention_mode, enc, keys, sv.h + (size_t)t * B * Hd, alpha, energy, enc_len,
                                               loc ? p->b_filter : nullptr, p->v_a, B, T, E, A, d->sharpening,
                                               d->sigmoid_smoothing, dctx_t, d_keys, dq_dst, kd < 1.f ? 0 : 1, g->v_a,
                                               loc ? g->b_filter : nullptr, prev_al, p->conv_filter, d->filter_width,
                                               p->w_filter, dal_in, dal_out, g->conv_filter, g->w_filter,
                                               att_ws, att_ws_bytes, stream_))) return rc;
      if (kd < 1.f)
        if ((rc = b2_dropout_rows(dc_tmp, Hd, dh_t, Hd, B, Hd, kd, d->dropout_seed, (uint64_t)t * B * Hd, Hd, 1, stream_))) return rc;
    }
    float* dz = dz_all + (size_t)t * B * 4 * Hd;
    float* dc_out = (t & 1) ? dc_a : dc_b;
    if ((rc = b2_lstm_cell_pointwise_backward(sv.z + (size_t)t * B * 4 * Hd, p->cell_bias, p->w_i_diag, p->w_f_diag,
                                              p->w_o_diag, sv.c + (size_t)t * B * Hd, dh_t, dc_in, B, Hd,
                                              d->forget_bias, d->cell_clip, dz, dc_out, stream_))) return rc;
    dc_in = dc_out;
    if ((rc = gemm_simt(0, 1, B, emb, 4 * Hd, 1.f, dz, 4 * Hd, k_emb, 4 * Hd, 0.f, demb_all + (size_t)t * B * emb, emb, nullptr, stream))) return rc;
    if (ke < 1.f)
      if ((rc = b2_dropout_rows(demb_all + (size_t)t * B * emb, emb, demb_all + (size_t)t * B * emb, emb, B, emb, ke,
                                d->dropout_seed + 1, (uint64_t)t * emb, (uint64_t)labels_ld * emb, 0, stream_))) return rc;
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

// =======================================================================================
// Beam search over the attention decoder
//   reference: models/attention/decoders/beam_search/beam_search_decoder.py:234-332
//   (beam_search_step), util.py:38-95 (mask_probs, normalize_score, choose_top_k), :14-26
//   (gather_tree_py).  The reference's wrapper (:25-232) is dead code limited to one utterance
//   (batch = beam); here every utterance of the batch carries its own beam: batch rows
//   r = u*W + w, the W rows of utterance u share its encoder states.
// Per step: the greedy iteration up to the logits, then beam_step_kernel (log-softmax, finished
// beams may only continue with <EOS>, length-normalised scores, top-W over W*C candidates -- only
// beam 0 at time 0 --, new log-probs / lengths / finished flags, parents + words appended to the
// history) and beam_gather_kernel (cell state, context, previous attention weights re-ordered by
// parent; next input = [embedding(word); context; h]).  At the end the histories are walked back
// (gather_tree).  Beam 0 is the best hypothesis (top-k output is sorted).
// =======================================================================================
namespace b2 {

int attention_step_forward_rows(int mode, const float* enc, const float* keys, const float* q,
                                const float* prev_alpha, const int32_t* enc_len,
                                const float* conv_filter, int filter_width,
                                const float* w_filter, const float* b_filter,
                                const float* v_a, int B, int T, int E, int A,
                                float sharpening_factor, int sigmoid_smoothing,
                                float* alpha, float* context, float* energy_out, int rows_per_utt,
                                b2_stream_t stream_);

constexpr int kBeamMaxW = 64;

struct BCand { float s; int idx; };
__device__ __forceinline__ bool bcand_better(const BCand& a, const BCand& b) {
  if (a.idx < 0) return false;
  if (b.idx < 0) return true;
  if (a.s > b.s) return true;
  if (a.s < b.s) return false;
  return a.idx < b.idx;                 // tf.nn.top_k: lower index first among equals
}

__global__ void __launch_bounds__(256)
beam_step_kernel(const float* __restrict__ logits, int W, int C, int t, int eos, float lpw, int use_penalty,
                 const float* __restrict__ lp_in, const int* __restrict__ fin_in, const int* __restrict__ len_in,
                 float* __restrict__ lp_out, int* __restrict__ fin_out, int* __restrict__ len_out,
                 float* __restrict__ score_out, int* __restrict__ parents, int* __restrict__ words,
                 float* __restrict__ S_tot, float* __restrict__ S_sc, int* __restrict__ done) {
  __shared__ BCand wbest[8];
  __shared__ int sel[kBeamMaxW];
  const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lg = logits + (size_t)u * W * C;
  float* tot = S_tot + (size_t)u * W * C;
  float* sc = S_sc + (size_t)u * W * C;
  const float FMIN = -3.402823466e38f;
  for (int w = warp; w < W; w += 8) {
    const float* row = lg + (size_t)w * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(row[c] - m);
    s = warp_sum(s);
    const float lse = m + __logf(s);
    const int fin = fin_in[u * W + w];
    const float base = lp_in[u * W + w];
    const int len = len_in[u * W + w];
    for (int c = lane; c < C; c += 32) {
      float pr = row[c] - lse;
      if (fin) pr = (c == eos) ? 0.f : FMIN;                       // mask_probs (util.py:38-68)
      const float tp = base + pr;
      const int nl = len + ((!fin && c != eos) ? 1 : 0);
      float score = tp;
      if (use_penalty) score = tp / (powf(5.f + (float)nl, lpw) / powf(6.f, lpw));   // util.py:71-95
      tot[w * C + c] = tp;
      sc[w * C + c] = (t == 0 && w > 0) ? __int_as_float(0x7fc00000) : score;             // time 0: beam 0 only
    }
  }
  __syncthreads();
  const int N = W * C;
  for (int r = 0; r < W; ++r) {
    BCand best; best.idx = -1; best.s = 0.f;
    for (int n = tid; n < N; n += 256) {
      const float v = sc[n];
      if (v != v) continue;
      BCand c; c.s = v; c.idx = n;
      if (bcand_better(c, best)) best = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      BCand other;
      other.s = __shfl_xor_sync(0xffffffffu, best.s, o);
      other.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
      if (bcand_better(other, best)) best = other;
    }
    if (lane == 0) wbest[warp] = best;
    __syncthreads();
    if (tid == 0) {
      BCand bb = wbest[0];
      for (int k = 1; k < 8; ++k) if (bcand_better(wbest[k], bb)) bb = wbest[k];
      sel[r] = bb.idx;
      if (bb.idx >= 0) { score_out[u * W + r] = bb.s; sc[bb.idx] = __int_as_float(0x7fc00000); }
    }
    __syncthreads();
  }
  if (tid < W) {
    const int idx = sel[tid];
    int parent = 0, word = eos, nf = 1, nl = 0;
    float nlp = FMIN;
    if (idx >= 0) {
      parent = idx / C; word = idx % C;
      nlp = tot[idx];
      nf = fin_in[u * W + parent] | (word == eos);
      nl = len_in[u * W + parent] + (nf ? 0 : 1);
    }
    lp_out[u * W + tid] = nlp; fin_out[u * W + tid] = nf; len_out[u * W + tid] = nl;
    parents[u * W + tid] = parent; words[u * W + tid] = word;
  }
  __syncthreads();
  if (tid == 0) {
    int all = 1;
    for (int w = 0; w < W; ++w) all &= fin_out[u * W + w];
    done[u] = all;
  }
}

// row r = (u, w): state and next input re-ordered by the chosen parent
__global__ void __launch_bounds__(256)
beam_gather_kernel(const int* __restrict__ parents, const int* __restrict__ words, int W, int Hd, int E, int T,
                   int emb, int C, const float* __restrict__ c_new, const float* __restrict__ h_new,
                   const float* __restrict__ ctx, const float* __restrict__ alpha,
                   const float* __restrict__ embedding, float* __restrict__ c_next, float* __restrict__ xh_next,
                   float* __restrict__ alpha_next) {
  const int r = blockIdx.x;
  const int u = r / W;
  const int src = u * W + parents[r];
  const int word = words[r];
  const int X = emb + E + Hd;
  float* xr = xh_next + (size_t)r * X;
  for (int i = threadIdx.x; i < Hd; i += 256) {
    c_next[(size_t)r * Hd + i] = c_new[(size_t)src * Hd + i];
    xr[emb + E + i] = h_new[(size_t)src * Hd + i];
  }
  for (int i = threadIdx.x; i < E; i += 256) xr[emb + i] = ctx[(size_t)src * E + i];
  for (int i = threadIdx.x; i < emb; i += 256)
    xr[i] = (word >= 0 && word < C) ? embedding[(size_t)word * emb + i] : 0.f;
  if (alpha_next)
    for (int i = threadIdx.x; i < T; i += 256) alpha_next[(size_t)r * T + i] = alpha[(size_t)src * T + i];
}

// xh0 / state of every beam row from the utterance's initial state
__global__ void __launch_bounds__(256)
beam_init_kernel(const float* __restrict__ embedding, int sos, const float* __restrict__ c0,
                 const float* __restrict__ h0, int W, int emb, int E, int Hd, int C, float* __restrict__ xh,
                 float* __restrict__ c_state, float* __restrict__ lp, int* __restrict__ fin, int* __restrict__ len) {
  const int r = blockIdx.x, u = r / W;
  const int X = emb + E + Hd;
  float* xr = xh + (size_t)r * X;
  for (int i = threadIdx.x; i < X; i += 256) {
    float v = 0.f;
    if (i < emb) v = (sos >= 0 && sos < C) ? embedding[(size_t)sos * emb + i] : 0.f;
    else if (i >= emb + E) v = h0[(size_t)u * Hd + (i - emb - E)];
    xr[i] = v;
  }
  for (int i = threadIdx.x; i < Hd; i += 256) c_state[(size_t)r * Hd + i] = c0[(size_t)u * Hd + i];
  if (threadIdx.x == 0) { lp[r] = 0.f; fin[r] = 0; len[r] = 0; }
}

// gather_tree (util.py:14-26): walk the parent pointers back from the last step
__global__ void beam_backtrack_kernel(const int* __restrict__ hist_words, const int* __restrict__ hist_parents,
                                      int steps, int R, int W, int L, int eos, int* __restrict__ out_ids) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int u = r / W;
  int cur = r % W;
  int* out = out_ids + (size_t)r * L;
  for (int t = steps - 1; t >= 0; --t) {
    out[t] = hist_words[(size_t)t * R + u * W + cur];
    cur = hist_parents[(size_t)t * R + u * W + cur];
  }
  for (int t = steps; t < L; ++t) out[t] = eos;
}

struct BeamWs {
  DecScratch s; float* xh2; float* c_state; float* c_next; float* alpha_prev;
  float* lp[2]; int* fin[2]; int* len[2]; int* parents; int* words; int* hist_w; int* hist_p;
  float* S_tot; float* S_sc; int* done;
};
static size_t beam_ws_layout(const b2_decoder_desc* d, int W, int L, void* base, BeamWs* w) {
  b2_decoder_desc dr = *d;
  dr.B = d->B * W;
  const size_t R = dr.B, X = (size_t)d->emb + d->E + d->Hd;
  size_t off = align_up(dec_scratch_layout(&dr, nullptr, nullptr), 256);
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o_c = take(R * d->Hd * 4), o_cn = take(R * d->Hd * 4), o_ap = take(R * d->T * 4);
  const size_t o_lp0 = take(R * 4), o_lp1 = take(R * 4), o_f0 = take(R * 4), o_f1 = take(R * 4);
  const size_t o_l0 = take(R * 4), o_l1 = take(R * 4), o_pa = take(R * 4), o_wd = take(R * 4);
  const size_t o_hw = take((size_t)L * R * 4), o_hp = take((size_t)L * R * 4);
  const size_t o_st = take(R * d->C * 4), o_ss = take(R * d->C * 4), o_dn = take((size_t)d->B * 4);
  (void)X;
  if (w) {
    char* p = (char*)base;
    dec_scratch_layout(&dr, base, &w->s);
    w->c_state = (float*)(p + o_c); w->c_next = (float*)(p + o_cn); w->alpha_prev = (float*)(p + o_ap);
    w->lp[0] = (float*)(p + o_lp0); w->lp[1] = (float*)(p + o_lp1);
    w->fin[0] = (int*)(p + o_f0); w->fin[1] = (int*)(p + o_f1);
    w->len[0] = (int*)(p + o_l0); w->len[1] = (int*)(p + o_l1);
    w->parents = (int*)(p + o_pa); w->words = (int*)(p + o_wd);
    w->hist_w = (int*)(p + o_hw); w->hist_p = (int*)(p + o_hp);
    w->S_tot = (float*)(p + o_st); w->S_sc = (float*)(p + o_ss); w->done = (int*)(p + o_dn);
  }
  return off;
}

}  // namespace b2

extern "C" size_t b2_attention_decoder_beam_workspace_bytes(const b2_decoder_desc* d, int beam_width, int max_steps) {
  if (!d || beam_width < 1 || max_steps < 1) return 0;
  return beam_ws_layout(d, beam_width, max_steps, nullptr, nullptr);
}

extern "C" int b2_attention_decoder_beam_search(const b2_decoder_desc* d, const b2_decoder_params* p,
                                                const float* enc, const float* keys, const int32_t* enc_len,
                                                const float* c0, const float* h0, int sos, int eos,
                                                int beam_width, float length_penalty_weight, int max_steps,
                                                int poll_every, int32_t* out_ids, int32_t* out_len,
                                                float* out_log_probs, float* out_scores, int32_t* steps_run,
                                                void* workspace, size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = dec_check(d);
  if (rc) return rc;
  B2_CHECK_ARG(p && enc && enc_len && c0 && h0 && out_ids && out_len && out_log_probs && out_scores && workspace,
               "b2_attention_decoder_beam_search: null pointer");
  B2_CHECK_ARG(beam_width >= 1 && beam_width <= kBeamMaxW, "b2_attention_decoder_beam_search: beam width %d not in [1,%d]",
               beam_width, kBeamMaxW);
  B2_CHECK_ARG(d->C >= beam_width, "b2_attention_decoder_beam_search: beam width %d exceeds the %d classes", beam_width, d->C);
  B2_CHECK_ARG(max_steps > 0, "b2_attention_decoder_beam_search: max_steps");
  const int W = beam_width, Bu = d->B, R = Bu * W, L = max_steps;
  const int T = d->T, E = d->E, Hd = d->Hd, A = d->A, emb = d->emb, C = d->C, X = emb + E + Hd;
  BeamWs w;
  if (workspace_bytes < beam_ws_layout(d, W, L, workspace, &w)) { set_error("b2_attention_decoder_beam_search: workspace too small"); return B2_ERR_WORKSPACE; }
  const bool loc = d->filter_width > 0;
  // the reference disables the penalty for weight None or 1 (util.py:88-91); 0 gives penalty 1
  const int use_penalty = (length_penalty_weight != 1.f && length_penalty_weight != 0.f) ? 1 : 0;
  float* xh = w.s.xh;
  beam_init_kernel<<<R, 256, 0, stream>>>(p->embedding, sos, c0, h0, W, emb, E, Hd, C, xh, w.c_state, w.lp[0],
                                          w.fin[0], w.len[0]);
  B2_LAUNCH_CHECK();
  const float* prev_alpha = nullptr;
  int* h_done = poll_every > 0 ? poll_buffer(Bu) : nullptr;
  float* c_state = w.c_state; float* c_next = w.c_next;
  int t = 0, cur = 0;
  for (; t < L; ++t) {
    float* q = d->query_projected ? w.s.q : w.s.h_new;
    float* xh_next = w.s.xh + (size_t)((t + 1) & 1) * R * X;
    if ((rc = gemm_simt(0, 0, R, 4 * Hd, X, 1.f, xh, X, p->cell_kernel, 4 * Hd, 0.f, w.s.z, 4 * Hd, nullptr, stream))) break;
    if ((rc = b2_lstm_cell_pointwise(w.s.z, p->cell_bias, p->w_i_diag, p->w_f_diag, p->w_o_diag, c_state, R, Hd,
                                     d->forget_bias, d->cell_clip, w.s.c_new, w.s.h_new, stream_))) break;
    if (d->query_projected)
      if ((rc = gemm_simt(0, 0, R, A, Hd, 1.f, w.s.h_new, Hd, p->w_query, A, 0.f, q, A, nullptr, stream))) break;
    if ((rc = attention_step_forward_rows(d->attention_mode, enc, keys, q, prev_alpha, enc_len,
                                          loc ? p->conv_filter : nullptr, d->filter_width, p->w_filter, p->b_filter,
                                          p->v_a, R, T, E, A, d->sharpening, d->sigmoid_smoothing, w.s.alpha, w.s.ctx,
                                          nullptr, W, stream_))) break;
    if ((rc = gemm_simt(0, 0, R, Hd, Hd, 1.f, w.s.h_new, Hd, p->w_av, Hd, 0.f, w.s.av, Hd, nullptr, stream))) break;
    if ((rc = gemm_simt(0, 0, R, Hd, E, 1.f, w.s.ctx, E, p->w_av + (size_t)Hd * Hd, Hd, 1.f, w.s.av, Hd, nullptr, stream))) break;
    if ((rc = b2_tanh_inplace(w.s.av, (int64_t)R * Hd, stream_))) break;
    if ((rc = gemm_simt(0, 0, R, C, Hd, 1.f, w.s.av, Hd, p->w_out, C, 0.f, w.s.logits, C, p->b_out, stream))) break;
    int* pa = w.hist_p + (size_t)t * R; int* wd = w.hist_w + (size_t)t * R;
    beam_step_kernel<<<Bu, 256, 0, stream>>>(w.s.logits, W, C, t, eos, length_penalty_weight, use_penalty,
                                             w.lp[cur], w.fin[cur], w.len[cur], w.lp[cur ^ 1], w.fin[cur ^ 1],
                                             w.len[cur ^ 1], out_scores, pa, wd, w.S_tot, w.S_sc, w.done);
    B2_LAUNCH_CHECK();
    beam_gather_kernel<<<R, 256, 0, stream>>>(pa, wd, W, Hd, E, T, emb, C, w.s.c_new, w.s.h_new, w.s.ctx, w.s.alpha,
                                              p->embedding, c_next, xh_next,
                                              d->feed_previous_attention ? w.alpha_prev : nullptr);
    B2_LAUNCH_CHECK();
    { float* tmp = c_state; c_state = c_next; c_next = tmp; }
    xh = xh_next;
    cur ^= 1;
    if (d->feed_previous_attention) prev_alpha = w.alpha_prev;
    if (h_done && (t + 1) % poll_every == 0 && t + 1 < L) {
      cudaMemcpyAsync(h_done, w.done, (size_t)Bu * sizeof(int), cudaMemcpyDeviceToHost, stream);
      cudaStreamSynchronize(stream);
      bool all = true;
      for (int b = 0; b < Bu; ++b) all = all && h_done[b];
      if (all) { ++t; break; }
    }
  }
  if (rc) return rc;
  const int steps = t < L ? t : L;
  beam_backtrack_kernel<<<cdiv(R, 128), 128, 0, stream>>>(w.hist_w, w.hist_p, steps, R, W, L, eos, out_ids);
  B2_LAUNCH_CHECK();
  B2_CUDA(cudaMemcpyAsync(out_len, w.len[cur], (size_t)R * 4, cudaMemcpyDeviceToDevice, stream));
  B2_CUDA(cudaMemcpyAsync(out_log_probs, w.lp[cur], (size_t)R * 4, cudaMemcpyDeviceToDevice, stream));
  if (steps_run) *steps_run = steps;
  return B2_OK;
}
