/*
 * b2asr.h -- C ABI of the B200-native acoustic-training hot path
 * (BLSTM encoder -> CTC forward/backward loss -> decoders -> clip/optimizer).
 *
 * This is the drop-in boundary for the hot path of
 * hirofumi0810/tensorflow_end2end_speech_recognition.  The reference has no
 * FFI of its own: every entry point below replaces one TensorFlow-1.x op that
 * the reference calls by name (file:line cited per function); the Python
 * host mirror (tensorflow_end2end_speech_recognition_b200/) binds them with
 * ctypes, see INTEGRATION.md.
 *
 * Conventions
 *  - plain C: POD + raw DEVICE pointers + sizes; no torch / C++ types.
 *  - every call is asynchronous on `stream` (a cudaStream_t); nothing
 *    synchronises, nothing allocates: scratch comes from the caller
 *    (`*_workspace_bytes` tells how much).
 *  - return value: 0 = ok, negative = B2_ERR_*; b2_last_error() returns a
 *    thread-local message for the last failing call.
 *  - all activations fp32 at the boundary (the reference is fp32 end to end);
 *    `precision` selects the arithmetic inside GEMM-shaped work.
 */
#ifndef B2ASR_H_
#define B2ASR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b2_stream_t; /* cudaStream_t */

enum {
  B2_OK = 0,
  B2_ERR_INVALID = -1,      /* bad argument (shape, null pointer, unsupported flag) */
  B2_ERR_WORKSPACE = -2,    /* workspace too small */
  B2_ERR_CUDA = -3,         /* CUDA runtime error (message in b2_last_error) */
  B2_ERR_UNSUPPORTED = -4   /* valid request the current build cannot serve */
};

/* arithmetic used inside GEMM-shaped work */
enum {
  B2_PREC_FP32 = 0,   /* CUDA-core fp32 FFMA: bit-for-bit deterministic, parity mode */
  B2_PREC_BF16 = 1    /* tcgen05 tensor cores: bf16 operands, fp32 accumulate (TMEM) */
};

int b2_version(void);
const char* b2_last_error(void);
/* 1 when the running device is sm_100 (tcgen05/TMA paths usable) */
int b2_device_is_sm100(void);
/* number of CUDA kernels this library has launched in this process (bench bookkeeping) */
unsigned long long b2_launch_count(void);

/* ------------------------------------------------------------------------ *
 * CTC loss + gradient                    replaces tf.nn.ctc_loss
 *   reference call: models/ctc/ctc.py:289-297 (ignore_longer=1) and
 *                   models/attention/joint_ctc_attention.py:308-316 (=0)
 * logits [T,B,C] time-major, unnormalised; labels in CSR form
 * (label_offsets[B+1] into labels_flat), the dense form of the SparseTensor
 * built by utils/io/labels/sparsetensor.py:12-39.  blank = C-1 in the
 * reference (ctc.py:101).  max_label_len = max_b L_b (host knows it).
 * loss[b] = -log p(l_b | x_b) (0 for skipped utterances, +inf when no
 * alignment exists); grad[t,b,c] = grad_scale * d loss[b] / d logits[t,b,c]
 * (softmax - occupancy; 0 for t >= seq_len[b]); grad may be NULL.
 * ------------------------------------------------------------------------ */
size_t b2_ctc_workspace_bytes(int T, int B, int C, int max_label_len);
int b2_ctc_loss_grad(const float* logits, const int32_t* labels_flat,
                     const int32_t* label_offsets, const int32_t* seq_len,
                     int T, int B, int C, int blank, int max_label_len,
                     int ignore_longer_outputs_than_inputs, float grad_scale,
                     float* loss, float* grad, void* workspace,
                     size_t workspace_bytes, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * CTC greedy decoder                     replaces tf.nn.ctc_greedy_decoder
 *   reference call: models/ctc/ctc.py:340-342; numpy twin
 *   models/ctc/decoders/greedy_decoder.py:19-50
 * logits [T,B,C] (or any per-frame monotone transform of the posteriors).
 * out_labels [B, T] int32 padded with -1; out_len [B].
 * ------------------------------------------------------------------------ */
int b2_ctc_greedy_decode(const float* logits, const int32_t* seq_len, int T,
                         int B, int C, int blank, int32_t* out_labels,
                         int32_t* out_len, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * CTC prefix beam search                 replaces BeamSearchDecoder.__call__
 *   reference: models/ctc/decoders/beam_search_decoder.py:53-152 (the decoder
 *   examples/librispeech/evaluation/eval_ctc.py:126-131 selects, width 20)
 * log_probs [B,T,C] natural-log posteriors (the reference takes np.log(probs)).
 * out_labels [B, T] padded -1; out_len [B]; out_score [B] = -log score.
 * ------------------------------------------------------------------------ */
size_t b2_ctc_beam_workspace_bytes(int T, int B, int C, int beam_width);
int b2_ctc_beam_decode(const float* log_probs, const int32_t* seq_len, int T,
                       int B, int C, int blank, int beam_width,
                       int32_t* out_labels, int32_t* out_len, float* out_score,
                       void* workspace, size_t workspace_bytes,
                       b2_stream_t stream);

/* TF-semantics CTC beam search      replaces tf.nn.ctc_beam_search_decoder as called at
 *   models/ctc/ctc.py:344-346 (beam_width, top_paths=1, merge_repeated=True).
 * logits [T,B,C] fp32 TIME-major raw scores (the op's own input), blank = C-1.
 * out_labels [B,T] padded -1, out_len [B], out_score [B] = log P_total of the best leaf
 * (logits minus the per-frame maximum, as TF 1.x accumulates it).  merge_repeated != 0 drops
 * a label that repeats its successor in the emitted path (TF's default). */
size_t b2_ctc_beam_tf_workspace_bytes(int T, int B, int C, int beam_width);
int b2_ctc_beam_decode_tf(const float* logits, const int32_t* seq_len, int T, int B, int C,
                          int blank, int beam_width, int merge_repeated,
                          int32_t* out_labels, int32_t* out_len, float* out_score,
                          void* workspace, size_t workspace_bytes, b2_stream_t stream);
/* softmax over the last axis of [rows, C]   (CTC.posteriors, ctc.py:354-380) */
int b2_softmax_rows(const float* x, float* y, int64_t rows, int C,
                    b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * GEMM with bias epilogue   replaces MatMul/BiasAdd under
 *   tf.contrib.layers.fully_connected (ctc.py:203,217) and the time-batched
 *   halves of LSTMBlockCell's [x,h].W (blstm.py:287-320)
 * C[M,N] (ldc) = alpha * op(A) . op(B) + beta * C + bias[N]
 *   transa = 0: A is [M,K] row-major (lda)   1: A is [K,M] row-major
 *   transb = 0: B is [K,N] row-major (ldb)   1: B is [N,K] row-major
 * fp32 in / fp32 out.  precision = B2_PREC_BF16 rounds operands to bf16 and
 * runs on tcgen05 (requires sm_100a; K, lda, ldb multiples of 8).
 * ------------------------------------------------------------------------ */
size_t b2_gemm_workspace_bytes(int M, int N, int K, int precision);
int b2_gemm(int transa, int transb, int M, int N, int K, float alpha,
            const float* A, int lda, const float* B, int ldb, float beta,
            float* C, int ldc, const float* bias, int precision,
            void* workspace, size_t workspace_bytes, b2_stream_t stream);

/* b2_gemm with a caller-kept bf16 shadow of A (same logical layout as A, row stride lda_lp elements,
 * 16-byte aligned, lda_lp % 8 == 0): the bf16 path then skips its fp32->bf16 pass over A -- e.g. the
 * encoder output [T*B, 2H] feeding the output layer (models/ctc/ctc.py:215-226), whose bf16 copy the
 * BLSTM layer already wrote (b2_blstm_reserve_y_lp).  A_lp == NULL or precision fp32: identical to b2_gemm. */
int b2_gemm_lp(int transa, int transb, int M, int N, int K, float alpha,
               const float* A, int lda, const void* A_lp, int lda_lp, const float* B, int ldb,
               float beta, float* C, int ldc, const float* bias, int precision,
               void* workspace, size_t workspace_bytes, b2_stream_t stream);
/* Same GEMM on operands that are already bf16 (uint16_t storage), no casts:
 *   a_mn = 0: A is [M rows][K contiguous] (pitch lda)   1: A is [K rows][M contiguous]
 *   b_mn = 0: B is [N rows][K contiguous] (pitch ldb)   1: B is [K rows][N contiguous]
 *   out_mode 0: C fp32 = alpha*A.B + bias   1: C fp32 += (split-K, fp32 atomics)
 *            2: C bf16 = alpha*A.B + bias
 * pitches must be multiples of 8 elements, base pointers 16-byte aligned (TMA). */
int b2_gemm_bf16(int a_mn, int b_mn, int M, int N, int K, float alpha,
                 const uint16_t* A, int lda, const uint16_t* B, int ldb, void* C,
                 int ldc, const float* bias, int out_mode, int k_splits,
                 b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Bidirectional LSTM layer              replaces LSTMBlockCell / LSTMCell /
 *   BasicLSTMCell under tf.nn.bidirectional_dynamic_rnn
 *   reference: models/encoders/core/blstm.py:258-332 (block), :187-255,
 *   :124-184; cell equations models/recurrent/layers/lstm.py:142-183
 * One call = one layer, both directions, all T steps.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t T, B, D_in, H;      /* max time, batch, input width, num_units    */
  int32_t use_peephole;       /* w_{i,f,o}_diag present                      */
  float forget_bias;          /* 1.0 in every reference call                 */
  float cell_clip;            /* <= 0: no clipping                           */
  float keep_prob;            /* DropoutWrapper(output_keep_prob); 1 = off   */
  uint64_t dropout_seed;      /* counter-hash seed for the mask              */
  int32_t precision;          /* B2_PREC_*                                   */
  int32_t need_backward;      /* 1: fill `reserve` for b2_blstm_layer_backward */
  int32_t num_proj;           /* LSTMCell num_proj (blstm.py:215-228): 0 = none; P > 0: the emitted and
                                 recurrent h is (o*tanh(c)) . projection [H,P]; kernel is [(D_in+P),4H],
                                 y [T,B,2P], final_state = c_fw [B,H], h_fw [B,P], c_bw [B,H], h_bw [B,P];
                                 fp32 CUDA-core path whatever `precision` says */
  /* backward-only hand-over of DropoutWrapper's mask between stacked layers (0 / 0 / 0 = off; only honoured where
   * b2_blstm_layer_path() == 1, B2_ERR_UNSUPPORTED elsewhere): */
  float dx_keep_prob;         /* in (0,1): the dx this call emits is multiplied by the output-dropout mask of the layer
                                 BELOW (its keep_prob / dropout_seed) inside the GEMM that produces it ...           */
  uint64_t dx_dropout_seed;
  int32_t dy_premasked;       /* ... and that layer's own backward call is told that its dy already carries the mask */
} b2_lstm_desc;

/* parameters of one direction, TF LSTMBlockCell layout:
 * kernel [(D_in+H), 4H] rows = [x; h], gate column blocks i, g(ci), f, o */
typedef struct {
  const float* kernel;
  const float* bias;        /* [4H] */
  const float* w_i_diag;    /* [H] or NULL */
  const float* w_f_diag;
  const float* w_o_diag;
  const float* projection;  /* [H, num_proj] or NULL */
} b2_lstm_params;

typedef struct {
  float* kernel;
  float* bias;
  float* w_i_diag;
  float* w_f_diag;
  float* w_o_diag;
  float* projection;
} b2_lstm_grads;

/* which implementation a layer of this shape runs on: 0 fp32 CUDA-core step kernels (or the bf16 hybrid with per-frame
 * recurrence), 1 cluster/TMEM tcgen05 recurrence (lstm_rec_tc.cu), 2 grid-resident wide-layer recurrence (lstm_wide.cu) */
int b2_blstm_layer_path(const b2_lstm_desc* d);
size_t b2_blstm_reserve_bytes(const b2_lstm_desc* d);
size_t b2_blstm_workspace_bytes(const b2_lstm_desc* d);

/* x [T,B,D_in] time-major; y [T,B,2H] = concat(fw, bw) after output dropout,
 * zero for t >= seq_len[b]; final_state [4,B,H] = c_fw, h_fw, c_bw, h_bw
 * (may be NULL).  x_lp: optional bf16 shadow of x ([T*B, D_in], pitch D_in) as
 * returned by b2_blstm_reserve_y_lp of the layer below (saves one cast pass in
 * the bf16 path); NULL is always valid. */
int b2_blstm_layer_forward(const b2_lstm_desc* d, const float* x, const void* x_lp,
                           const int32_t* seq_len, const b2_lstm_params* fw,
                           const b2_lstm_params* bw, float* y,
                           float* final_state, void* reserve, void* workspace,
                           size_t workspace_bytes, b2_stream_t stream);
/* bf16 copy of the layer output kept inside `reserve` (bf16 path), else NULL */
const void* b2_blstm_reserve_y_lp(const b2_lstm_desc* d, const void* reserve);

/* dy [T,B,2H]; dx [T,B,D_in] (NULL for the first layer); gradients are
 * ACCUMULATED into g_fw / g_bw (caller zeroes them once per step). */
int b2_blstm_layer_backward(const b2_lstm_desc* d, const float* x, const void* x_lp,
                            const int32_t* seq_len, const b2_lstm_params* fw,
                            const b2_lstm_params* bw, const float* dy,
                            const void* reserve, float* dx,
                            const b2_lstm_grads* g_fw, const b2_lstm_grads* g_bw,
                            void* workspace, size_t workspace_bytes,
                            b2_stream_t stream);

/* Same, with the gradient of the layer's final state d_final_state [4,B,H] =
 * d(c_fw, h_fw, c_bw, h_bw) (NULL: none) -- the path by which the attention model's bridge
 * (models/attention/bridge.py:128-151) back-propagates into the encoder. */
int b2_blstm_layer_backward_ex(const b2_lstm_desc* d, const float* x, const void* x_lp,
                               const int32_t* seq_len, const b2_lstm_params* fw,
                               const b2_lstm_params* bw, const float* dy,
                               const float* d_final_state, const void* reserve, float* dx,
                               const b2_lstm_grads* g_fw, const b2_lstm_grads* g_bw,
                               void* workspace, size_t workspace_bytes,
                               b2_stream_t stream);

/* Measurement aid: CUDA-event timers around the persistent recurrence kernels of the most
 * recent bf16 layer forward / backward (caller synchronises before reading). */
void b2_blstm_profile_enable(int on);
int b2_blstm_profile_last_ms(float* fwd_ms, float* bwd_ms);

/* In the bf16 path the weight-gradient GEMMs of a layer run on an internal low-priority
 * side stream (they overlap the next layer's BPTT recurrence).  Call this once after the
 * last b2_blstm_layer_backward of a step: it makes `stream` wait for them (no host sync). */
int b2_blstm_backward_join(b2_stream_t stream);
/* `stream` waits for the side-stream work ENQUEUED so far -- after the backward call of layer l that is the
 * weight gradients of layer l+1 (those of layer l itself are launched next to layer l-1's recurrence, or by the
 * join).  The data-parallel step reduces layer l+1's gradient bucket behind this wait while BPTT continues
 * (replaces the single synchronisation point of utils/training/multi_gpu.py:13-48). */
int b2_blstm_backward_side_wait(b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Input pipeline on the device: frame stacking / skipping + splicing + zero padding
 *   reference: utils/io/inputs/frame_stacking.py:14-85 (stack_frame),
 *   utils/io/inputs/splicing.py:9-73 (do_splice), applied per utterance before padding by
 *   utils/dataset/ctc.py:120-160.  Bit-exact gather (golden vectors from the reference's own
 *   numpy functions: tests/golden/input_pipeline.npz).
 * raw [B,Traw,D] zero-padded features, raw_len [B] -> out [B,Tout,Dout] (zero past each
 * utterance's new length), out_len [B] = raw_len (num_stack == 1) or ceil(raw_len/num_skip).
 * Dout = b2_stack_splice_out_dim(D, num_stack, splice).
 * ------------------------------------------------------------------------ */
int b2_stack_splice_out_dim(int D, int num_stack, int splice);
int b2_stack_splice(const float* raw, const int32_t* raw_len, int B, int Traw, int D,
                    int num_stack, int num_skip, int splice, int Tout, float* out,
                    int32_t* out_len, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * VGG front-end of the VGG-BLSTM encoder
 *   reference: models/encoders/core/vgg_blstm.py:93-177 (VGGBLSTMEncoder.__call__ up to the
 *   BLSTM), models/encoders/core/cnn_util.py:52-84 (conv_layer), :13-29 (max_pool)
 * x [N=B*T, H=num_channels, W=splice*num_stack, 3] (the reference's reshape of [B,T,D]) ->
 * conv3x3(3->64) ReLU, dropout, conv(64->64) ReLU, max_pool 2x2/2 SAME, dropout,
 * conv(64->128) ReLU, dropout, conv(128->128) ReLU, max_pool, dropout, flatten (h,w,c),
 * fully connected 256 + ReLU, dropout -> out [N,256].
 * Filters are TF layout [3,3,C_in,C_out]; fc_w [b2_vgg_output_size(d), 256].
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t N, H, W;            /* frames, num_channels, splice*num_stack      */
  float keep_prob;            /* tf.nn.dropout keep probability; 1 = off      */
  uint64_t dropout_seed;      /* counter-hash seed (sites use seed+1..seed+5) */
  int32_t precision;          /* B2_PREC_FP32: CUDA-core fp32 GEMMs; B2_PREC_BF16: the 64/128-channel
                                 convolutions and the bridge FC on tcgen05 (bf16 operands, fp32 accumulate) */
} b2_vgg_desc;
typedef struct {
  const float* conv_w[4];     /* VGG1/conv1, VGG1/conv2, VGG2/conv1, VGG2/conv2 */
  const float* conv_b[4];
  const float* fc_w;          /* bridge/weights */
  const float* fc_b;          /* bridge/biases  */
} b2_vgg_params;
typedef struct {
  float* conv_w[4];
  float* conv_b[4];
  float* fc_w;
  float* fc_b;
} b2_vgg_grads;
size_t b2_vgg_reserve_bytes(const b2_vgg_desc* d);
size_t b2_vgg_workspace_bytes(const b2_vgg_desc* d);
int b2_vgg_output_size(const b2_vgg_desc* d);   /* flattened width feeding the FC */
int b2_vgg_frontend_forward(const b2_vgg_desc* d, const float* x, const b2_vgg_params* p,
                            float* out, void* reserve, void* workspace,
                            size_t workspace_bytes, b2_stream_t stream);
/* gradients are ACCUMULATED into g; there is no input gradient (x is the feature matrix) */
int b2_vgg_frontend_backward(const b2_vgg_desc* d, const b2_vgg_params* p, const float* d_out,
                             const void* reserve, const b2_vgg_grads* g, void* workspace,
                             size_t workspace_bytes, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * The whole attention decoder loop behind one call per direction
 *   reference: tf.while_loop over AttentionDecoder.step -- dynamic_decoder.py:148-212,
 *   attention_decoder.py:142-295; helpers attention_seq2seq.py:440-446 (TrainingHelper),
 *   :486-494 (GreedyEmbeddingHelper)
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t B, T, E, Hd, A, emb, C;   /* batch, encoder frames, encoder width, decoder units, query
                                       width, embedding dim, classes (incl. <SOS>, <EOS>)     */
  int32_t attention_mode;           /* 0 additive, 1 multiplicative (b2_attention_step_forward) */
  int32_t query_projected;          /* 1: q = h . w_query [Hd,A]; 0: q = h (A == Hd)          */
  int32_t filter_width;             /* location term: conv filter taps, 0 = none              */
  float sharpening;
  int32_t sigmoid_smoothing;
  float forget_bias, cell_clip;     /* LSTMBlockCell; cell_clip <= 0: none                     */
  int32_t feed_previous_attention;  /* 0 = the reference's behaviour (zeros), inference only   */
  float keep_prob_decoder;          /* DropoutWrapper(output_keep_prob) on the decoder cell output; training only */
  float keep_prob_embedding;        /* tf.nn.dropout on the embedded labels (attention_seq2seq.py:438-439)        */
  uint64_t dropout_seed;            /* cell output mask index (t*B+b)*Hd+i at seed; embedding mask index
                                       (b*labels_ld+pos)*emb+i at seed+1                                           */
} b2_decoder_desc;
typedef struct {
  const float* cell_kernel;         /* [emb+E+Hd, 4Hd], gate blocks i,g,f,o                     */
  const float* cell_bias;           /* [4Hd]                                                   */
  const float* w_i_diag; const float* w_f_diag; const float* w_o_diag;   /* [Hd] or all NULL   */
  const float* w_query;             /* [Hd,A] or NULL                                          */
  const float* conv_filter;         /* [filter_width,10] or NULL                               */
  const float* w_filter;            /* [10,A]                                                  */
  const float* b_filter;            /* [A]                                                     */
  const float* v_a;                 /* [A] or NULL                                             */
  const float* w_av;                /* attentional_vector/weights [Hd+E, Hd]                   */
  const float* w_out;               /* output_layer/weights [Hd, C]                            */
  const float* b_out;               /* [C]                                                     */
  const float* embedding;           /* W_embedding [C, emb]                                    */
} b2_decoder_params;
typedef struct {                    /* same fields, gradients are ACCUMULATED                  */
  float* cell_kernel; float* cell_bias; float* w_i_diag; float* w_f_diag; float* w_o_diag;
  float* w_query; float* conv_filter; float* w_filter; float* b_filter; float* v_a;
  float* w_av; float* w_out; float* b_out; float* embedding;
} b2_decoder_grads;
size_t b2_attention_decoder_reserve_bytes(const b2_decoder_desc* d, int max_steps);
size_t b2_attention_decoder_workspace_bytes(const b2_decoder_desc* d, int max_steps);
/* labels == NULL: greedy decoding from <SOS> until every row emitted <EOS> or max_steps (the
 * finished flags are polled on the host every poll_every steps, 0 = never); labels [B,labels_ld]
 * + dec_len [B]: teacher forcing on labels[:, :-1].  reserve != NULL (teacher forcing only): keep
 * what b2_attention_decoder_backward needs.  keys: the hoisted key projection [B,T,A] (NULL for
 * `location`).  Outputs are batch-major [B,max_steps,*] with zeros past each row's finish;
 * c_state/h_state [B,Hd] = final state; *steps_run = iterations executed. */
int b2_attention_decoder_forward(const b2_decoder_desc* d, const b2_decoder_params* p,
                                 const float* enc, const float* keys, const int32_t* enc_len,
                                 const float* c0, const float* h0, const int32_t* labels,
                                 int labels_ld, const int32_t* dec_len, int sos, int eos,
                                 int max_steps, int poll_every, void* reserve,
                                 float* out_logits, int32_t* out_ids, float* out_av,
                                 float* out_alpha, float* out_ctx, float* c_state,
                                 float* h_state, int32_t* finished, int32_t* steps_run,
                                 void* workspace, size_t workspace_bytes, b2_stream_t stream);
/* dlogits_tm [steps,B,C] time-major.  Accumulates parameter gradients into g, d_keys [B,T,A]
 * (NULL for `location`; for luong_dot pass d_enc) and d_enc [B,T,E] (through the context only --
 * the key projection's share is the caller's GEMM); writes dc0, dh0 [B,Hd]. */
int b2_attention_decoder_backward(const b2_decoder_desc* d, const b2_decoder_params* p,
                                  const float* enc, const float* keys, const int32_t* enc_len,
                                  const int32_t* labels, int labels_ld, int steps,
                                  const void* reserve, const float* dlogits_tm,
                                  const b2_decoder_grads* g, float* d_keys, float* d_enc,
                                  float* dc0, float* dh0, void* workspace,
                                  size_t workspace_bytes, b2_stream_t stream);

/* Beam search over the attention decoder (models/attention/decoders/beam_search/
 * beam_search_decoder.py:234-332 beam_search_step; util.py:38-95 mask_probs / normalize_score /
 * choose_top_k; :14-26 gather_tree).  d->B utterances, each with its own beam of beam_width
 * hypotheses (batch rows u*W + w share utterance u's enc / keys / enc_len); c0, h0 [B,Hd].
 * Finished beams may only continue with <EOS>; scores = log_prob / ((5+len)^w / 6^w) unless the
 * weight is 0 or 1 (the reference disables the penalty for 1); at step 0 only beam 0 is expanded.
 * Stops when every beam of every utterance is finished (polled every poll_every steps, 0 = never)
 * or after max_steps.  out_ids [B,W,max_steps] (<EOS> past the end), out_len / out_log_probs /
 * out_scores [B,W]; beam 0 is the best hypothesis. */
size_t b2_attention_decoder_beam_workspace_bytes(const b2_decoder_desc* d, int beam_width, int max_steps);
int b2_attention_decoder_beam_search(const b2_decoder_desc* d, const b2_decoder_params* p,
                                     const float* enc, const float* keys, const int32_t* enc_len,
                                     const float* c0, const float* h0, int sos, int eos,
                                     int beam_width, float length_penalty_weight, int max_steps,
                                     int poll_every, int32_t* out_ids, int32_t* out_len,
                                     float* out_log_probs, float* out_scores, int32_t* steps_run,
                                     void* workspace, size_t workspace_bytes, b2_stream_t stream);

/* Levenshtein distance of B (hypothesis, reference) label-sequence pairs (tf.edit_distance as
 * compute_ler uses it, models/ctc/ctc.py:382-398): flat label arrays + [B+1] offsets, dist [B].
 * The caller divides by the reference length (normalize=True) and averages. */
int b2_edit_distance(const int32_t* hyp, const int32_t* hyp_offsets, const int32_t* ref,
                     const int32_t* ref_offsets, int B, int max_ref_len, int32_t* dist,
                     b2_stream_t stream);

/* x = dropout(relu(x)) in place (element i keeps with the counter hash of (seed, i)) and its
 * backward d_in = d_out * (out > 0 ? 1/keep_prob : 0): the activation of the VGG bridge layer and
 * of the CTC bottleneck layer (models/ctc/ctc.py:200-213). */
int b2_relu_dropout_forward(float* x, int64_t n, float keep_prob, uint64_t seed, b2_stream_t stream);
int b2_relu_dropout_backward(const float* d_out, const float* out, int64_t n, float keep_prob,
                             float* d_in, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Attention step (energy + masked softmax + context)   replaces
 *   AttentionLayer.__call__, models/attention/decoders/attention_layer.py:45-347
 * One decoder step over all T encoder states; the key projection is hoisted
 * out of the step (b2_gemm once per batch):
 *   mode 0 additive:       e_t = sum_a v_a*tanh(keys[t,a] + q[a] + loc[t,a])
 *                          (bahdanau_content, hybrid, location, luong_concat)
 *   mode 1 multiplicative: e_t = sum_a keys[t,a]*q[a]
 *                          (dot_product, luong_dot, luong_general)
 *   loc = conv1d_SAME(prev_alpha, conv_filter[filter_width,10]) . w_filter[10,A]
 *         + b_filter (NULL conv_filter = no location term; keys may be NULL
 *         for pure `location`; prev_alpha NULL with a location term = all-zero
 *         previous weights, the convolution is skipped)
 * enc [B,T,E], keys [B,T,A], q [B,A], prev_alpha [B,T], enc_len [B];
 * energies of t >= enc_len are float32.min, then *sharpening_factor, then
 * softmax (or sigmoid / sum when sigmoid_smoothing).  alpha [B,T], context [B,E].
 * energy_out (may be NULL): the sharpened energies [B,T] of the valid frames, kept for
 * the backward pass of the sigmoid-smoothing form.
 * ------------------------------------------------------------------------ */
int b2_attention_step_forward(int mode, const float* enc, const float* keys,
                              const float* q, const float* prev_alpha,
                              const int32_t* enc_len, const float* conv_filter,
                              int filter_width, const float* w_filter,
                              const float* b_filter, const float* v_a, int B,
                              int T, int E, int A, float sharpening_factor,
                              int sigmoid_smoothing, float* alpha,
                              float* context, float* energy_out,
                              b2_stream_t stream);

/* Backward of one attention step given d(context) [B,E] of that decoder step
 * (attention_layer.py:45-113 differentiated).  Writes dq [B,A] (adds to it when
 * dq_accumulate != 0); accumulates (+=) into
 * d_keys [B,T,A] (NULL: no key term), dv [A] (NULL for multiplicative scores) and
 * db_filter [A] (NULL: no location term).  The location term is the constant b_filter
 * (the reference's decoder always feeds zero previous weights), so conv filter and W_filter
 * receive no gradient.  d(enc) through the context, alpha_t (x) dctx_t, is NOT added here:
 * it is a rank-(number of steps) update per utterance done by one b2_gemm after the loop.
 * alpha/energy: the forward outputs of this step.  workspace from
 * b2_attention_step_backward_workspace_bytes(B, T). */
size_t b2_attention_step_backward_workspace_bytes(int B, int T);
int b2_attention_step_backward(int mode, const float* enc, const float* keys, const float* q,
                               const float* alpha, const float* energy,
                               const int32_t* enc_len, const float* b_filter,
                               const float* v_a, int B, int T, int E, int A,
                               float sharpening_factor, int sigmoid_smoothing,
                               const float* dctx, float* d_keys, float* dq,
                               int dq_accumulate, float* dv, float* db_filter, void* workspace, size_t workspace_bytes,
                               b2_stream_t stream);

/* Same with the location term fed by REAL previous weights (feed_previous_attention; the intended
 * behaviour the reference's while_loop never reaches): prev_alpha [B,T] != NULL makes the
 * kernel recompute conv1d(prev_alpha, conv_filter) . w_filter, accumulate d_conv_filter [filter_width,10]
 * and d_w_filter [10,A], and write d_prev_alpha [B,T] (gradient wrt the previous step's weights);
 * dalpha_ext [B,T] (may be NULL) is the gradient wrt THIS step's weights that arrives from the next
 * step's location term. */
int b2_attention_step_backward_loc(int mode, const float* enc, const float* keys, const float* q,
                                   const float* alpha, const float* energy,
                                   const int32_t* enc_len, const float* b_filter,
                                   const float* v_a, int B, int T, int E, int A,
                                   float sharpening_factor, int sigmoid_smoothing,
                                   const float* dctx, float* d_keys, float* dq,
                                   int dq_accumulate, float* dv, float* db_filter,
                                   const float* prev_alpha, const float* conv_filter, int filter_width,
                                   const float* w_filter, const float* dalpha_ext,
                                   float* d_prev_alpha, float* d_conv_filter, float* d_w_filter,
                                   void* workspace, size_t workspace_bytes, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Attention decoder, training side
 *   reference: attention_seq2seq.py:579-664 (compute_loss), :413-459 (_decode_train)
 * ------------------------------------------------------------------------ */
/* tf.contrib.seq2seq.sequence_loss(average_across_timesteps=True, average_across_batch=True)
 * with weights = sequence_mask(lengths, L) on logits/temperature [B,L,C] and targets
 * [B, >=L] (row stride targets_ld):  rowloss[b*L+t] = w*xent (the caller divides their sum by
 * sum(w) + 1e-12);  dlogits (may be NULL) = grad_scale * d(loss)/d(logits). */
int b2_sequence_loss(const float* logits, const int32_t* targets, int targets_ld,
                     const int32_t* lengths, int B, int L, int C, float temperature,
                     float grad_scale, float* rowloss, float* dlogits,
                     b2_stream_t stream);
int b2_tanh_backward(const float* dy, const float* y, float* dx, int64_t n, b2_stream_t stream);
/* backward of b2_lstm_cell_pointwise: dz [B,4H] (gate blocks i,g,f,o), dc_prev [B,H];
 * dc_in may be NULL (no gradient from the next step) */
int b2_lstm_cell_pointwise_backward(const float* z, const float* bias, const float* w_i_diag,
                                    const float* w_f_diag, const float* w_o_diag,
                                    const float* c_prev, const float* dh, const float* dc_in,
                                    int B, int H, float forget_bias, float cell_clip,
                                    float* dz, float* dc_prev, b2_stream_t stream);
/* peephole gradients over all decoder steps: dz [steps*B,4H]; c_all [(steps+1)*B,H] where
 * block s is the cell state before step s (block 0 = initial state).  Accumulates. */
int b2_decoder_peephole_grad(const float* dz, const float* c_all, int steps, int B, int H,
                             float* dw_i, float* dw_f, float* dw_o, b2_stream_t stream);
/* tf.nn.dropout on a row-strided matrix with explicit element numbering: (r, c) is kept iff the counter
 * hash of (seed, idx_base + r*idx_row_stride + c) says so; y = mask*x/keep_prob, or y += ... when add != 0
 * (x == y allowed).  Used for the decoder-cell output and the label embedding of the attention model. */
int b2_dropout_rows(const float* x, int ldx, float* y, int ldy, int64_t rows, int cols, float keep_prob,
                    uint64_t seed, uint64_t idx_base, uint64_t idx_row_stride, int add, b2_stream_t stream);
/* dW[ids[r], :D] += dx[r, :D] for r < rows (dx row stride ldx); ids outside [0,V) skipped */
int b2_embedding_grad(const float* dx, int ldx, const int32_t* ids, int64_t rows, int D, int V,
                      float* dW, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Attention decoder step, forward (greedy inference path)
 *   reference: models/attention/decoders/attention_decoder.py:256-295 (step),
 *   :170-211 (_compute_output), attention_seq2seq.py:352-363 (LSTMBlockCell),
 *   :490-494 (GreedyEmbeddingHelper)
 * ------------------------------------------------------------------------ */
/* gate math of one LSTMBlockCell step on z[B,4H] = [x,h].W (gate blocks i,g,f,o); bias and
 * peepholes may be NULL; cell_clip <= 0: none */
int b2_lstm_cell_pointwise(const float* z, const float* bias, const float* w_i_diag,
                           const float* w_f_diag, const float* w_o_diag,
                           const float* c_prev, int B, int H, float forget_bias,
                           float cell_clip, float* c_out, float* h_out,
                           b2_stream_t stream);
int b2_tanh_inplace(float* x, int64_t n, b2_stream_t stream);
/* Tail of one dynamic_decode iteration (dynamic_decoder.py:148-196 with impute_finished, the
 * input-feeding next_inputs of attention_decoder.py:221-238 and the helper's finished rule):
 * writes the step's outputs into slot t of the batch-major [B,L,*] output arrays (zeros for
 * finished rows), copies the state through for finished rows, builds the next cell input
 * xh[B, emb+E+Hd] = [embedding(next id) ; ctx ; h_state] and updates finished[B] in place.
 * labels == NULL: greedy helper (next id = ids[b], finished when it equals eos);
 * labels [B,labels_ld] + dec_len[B]: training helper on labels[:, :-1] (finished when
 * t + 1 >= dec_len[b]).  max_iter > 0: everything is finished once t + 1 >= max_iter. */
int b2_decoder_step_emit(int B, int C, int Hd, int E, int T, int emb_dim, int t, int L,
                         const float* logits, const int32_t* ids, const float* av,
                         const float* alpha, const float* ctx, const float* c_new,
                         const float* h_new, float* c_state, float* h_state,
                         int32_t* finished, const float* embedding,
                         const int32_t* labels, int labels_ld, const int32_t* dec_len,
                         int eos, int max_iter, float* xh, float* out_logits,
                         int32_t* out_ids, float* out_av, float* out_alpha,
                         float* out_ctx, b2_stream_t stream);
/* out[r] = argmax_c x[r, c], first index on ties */
int b2_argmax_rows(const float* x, int64_t rows, int C, int32_t* out, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Small data-movement helpers of the step
 * ------------------------------------------------------------------------ */
/* [B,T,D] -> [T,B,D]  (tf.transpose at blstm.py:279) */
int b2_transpose_01(const float* x, float* y, int d0, int d1, int d2,
                    b2_stream_t stream);
/* column sums: out[N] (+)= sum_m X[m, n]   (bias gradient) */
int b2_colsum(const float* X, int64_t M, int N, int ldx, float* out,
              int accumulate, b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Gradient clipping + optimizers          replaces tf.clip_by_norm and
 *   tf.train.*Optimizer      reference: models/model_base.py:12-20,68-95,135-166
 * Multi-tensor: n tensors described by device arrays of pointers/sizes.
 * ------------------------------------------------------------------------ */
enum {
  B2_OPT_SGD = 0, B2_OPT_MOMENTUM = 1, B2_OPT_NESTEROV = 2, B2_OPT_ADAGRAD = 3,
  B2_OPT_ADADELTA = 4, B2_OPT_ADAM = 5, B2_OPT_RMSPROP = 6
};
/* per-tensor tf.clip_by_norm, in place; norms [n] is scratch/diagnostic out.
 * post_scale is applied after clipping (1/num_towers for the tower mean). */
int b2_clip_by_norm_multi(float* const* grads, const int64_t* sizes, int n,
                          float clip_norm, float post_scale, float* norms,
                          b2_stream_t stream);
/* y_k += alpha * x_k over n tensors (weight-decay gradient wd*w, ctc.py:280-286) */
int b2_axpy_multi(float* const* xs, float* const* ys, const int64_t* sizes, int n,
                  float alpha, b2_stream_t stream);
/* Tower mean   replaces average_gradients, utils/training/multi_gpu.py:13-48 (tf.concat + reduce_mean
 * per variable on /cpu:0): dst[i] = mean_k srcs[k][i] over n_src <= 16 device buffers of n floats.
 * srcs is a HOST array of device pointers; dst may alias srcs[0]. */
int b2_tower_mean(const float* const* srcs, int n_src, float* dst, int64_t n,
                  b2_stream_t stream);
/* TF-1.x update rules; state0/state1 per tensor (may be NULL
 * where the rule needs none); step = 1-based global step (Adam). */
int b2_optimizer_step_multi(int kind, float* const* params,
                            float* const* grads, float* const* state0,
                            float* const* state1, const int64_t* sizes, int n,
                            float learning_rate, int64_t step,
                            b2_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Gradient exchange of the data-parallel step   replaces average_gradients,
 *   utils/training/multi_gpu.py:13-48, as called at
 *   examples/librispeech/training/train_ctc.py:143 (the synchronisation point
 *   of the towers).  One communicator rank per GPU; NCCL (>= 2.10) is bound at
 *   run time, b2_comm_available() returns its version or 0.
 *   Bootstrap: rank 0 calls b2_comm_get_unique_id (128 bytes), the caller ships
 *   the id to the other ranks (any side channel), every rank calls
 *   b2_comm_init_rank with its own device current.
 * ------------------------------------------------------------------------ */
typedef void* b2_comm_t;
#define B2_COMM_ID_BYTES 128
int b2_comm_available(void);
int b2_comm_get_unique_id(void* id_out);
int b2_comm_init_rank(b2_comm_t* comm_out, int nranks, const void* id, int rank);
/* one communicator per local device of a single process (in-graph towers) */
int b2_comm_init_all(b2_comm_t* comms_out, int ndev, const int* devices);
int b2_comm_size(b2_comm_t comm);
int b2_comm_destroy(b2_comm_t comm);
/* in-place mean over ranks of n_buckets device buffers (HOST arrays of pointers
 * and element counts), one NCCL group call on `stream` */
int b2_allreduce_mean(b2_comm_t comm, float* const* buckets, const int64_t* sizes,
                      int n_buckets, b2_stream_t stream);
/* single-process form: buffer k (n floats) lives on the device of comms[k] and
 * is reduced on streams[k]; all ndev calls inside one group */
int b2_allreduce_mean_local(const b2_comm_t* comms, float* const* buffers, int64_t n,
                            int ndev, const b2_stream_t* streams);

/* CRC-32C (Castagnoli) of a host buffer, continuing from `crc` (0 to start): the checksum of TensorFlow's checkpoint
 * bundles, used by utils/io/tf_checkpoint.py (the Saver interop of examples/timit/evaluation/eval_ctc.py:70-88). */
uint32_t b2_crc32c(uint32_t crc, const void* data, size_t n);

/* ------------------------------------------------------------------ GRU layers
 * tf.contrib.rnn.GRUCell under bidirectional_dynamic_rnn (models/encoders/core/gru.py:128-160, BGRUEncoder) or
 * MultiRNNCell + dynamic_rnn (gru.py:52-73, GRUEncoder: bind an all-zero second direction):
 *   [r, u] = sigmoid([x, h] . gates_kernel + gates_bias),  c = tanh([x, r*h] . cand_kernel + cand_bias),
 *   h' = u*h + (1-u)*c;  sequence_length semantics and output dropout as in the LSTM layers.  fp32.
 *   x [T,B,D_in] time-major; y [T,B,2H] = concat(fw, bw); final_state [2,B,H] = h_fw, h_bw (may be NULL);
 *   gradients are ACCUMULATED into g_fw / g_bw; dx may be NULL. */
typedef struct {
  int32_t T, B, D_in, H;
  float keep_prob;            /* DropoutWrapper(output_keep_prob); 1 = off */
  uint64_t dropout_seed;
  int32_t need_backward;
} b2_gru_desc;
typedef struct {
  const float* gates_kernel;  /* [(D_in+H), 2H] rows = [x; h], columns r | u */
  const float* gates_bias;    /* [2H] (TF initialises it to 1) */
  const float* cand_kernel;   /* [(D_in+H), H]  rows = [x; r*h] */
  const float* cand_bias;     /* [H] */
} b2_gru_params;
typedef struct {
  float* gates_kernel; float* gates_bias; float* cand_kernel; float* cand_bias;
} b2_gru_grads;
size_t b2_bgru_reserve_bytes(const b2_gru_desc* d);
size_t b2_bgru_workspace_bytes(const b2_gru_desc* d);
int b2_bgru_layer_forward(const b2_gru_desc* d, const float* x, const int32_t* seq_len, const b2_gru_params* fw,
                          const b2_gru_params* bw, float* y, float* final_state, void* reserve, void* workspace,
                          size_t workspace_bytes, b2_stream_t stream);
int b2_bgru_layer_backward(const b2_gru_desc* d, const float* x, const int32_t* seq_len, const b2_gru_params* fw,
                           const b2_gru_params* bw, const float* dy, const void* reserve, float* dx,
                           const b2_gru_grads* g_fw, const b2_gru_grads* g_bw, void* workspace,
                           size_t workspace_bytes, b2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2ASR_H_ */
