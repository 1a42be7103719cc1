// Training-side kernels of the attention decoder: sequence loss + gradient, and the backward
// pieces of one decoder step.  All HBM/latency-bound elementwise or row-reduction work.
//
//   b2_sequence_loss              tf.contrib.seq2seq.sequence_loss with sequence_mask weights,
//                                 logits / temperature          (attention_seq2seq.py:270,619-636)
//   b2_tanh_backward              d/dx tanh through the attentional vector (attention_decoder.py:189-196)
//   b2_lstm_cell_pointwise_backward   LSTMBlockCell gate math backward (attention_seq2seq.py:352-363,
//                                 equations models/recurrent/layers/lstm.py:142-183)
//   b2_decoder_peephole_grad      peephole gradients over all decoder steps
//   b2_embedding_grad             scatter-add of d(embedded labels) into W_embedding
//                                 (tf.nn.embedding_lookup backward, attention_seq2seq.py:436-437)
#include "common.cuh"

namespace b2 {

// one warp per (b, t) row.  rowloss[b*L+t] = w * xent ; dlogits = w * (softmax - onehot) * scale
__global__ void __launch_bounds__(256)
sequence_loss_kernel(const float* __restrict__ logits, const int* __restrict__ targets, int ldt,
                     const int* __restrict__ lengths, int B, int L, int C, float inv_temp,
                     float grad_scale_num, float* __restrict__ rowloss, float* __restrict__ dlogits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= (int64_t)B * L) return;
  const int b = (int)(row / L), t = (int)(row % L);
  const bool on = t < lengths[b];
  float* dr = dlogits ? dlogits + row * C : nullptr;
  if (!on) {
    if (lane == 0) rowloss[row] = 0.f;
    if (dr) for (int c = lane; c < C; c += 32) dr[c] = 0.f;
    return;
  }
  // sum of weights = sum_b min(len_b, L) (every warp recomputes it: B is small)
  float wsum = 0.f;
  for (int i = lane; i < B; i += 32) wsum += (float)max(0, min(lengths[i], L));
  wsum = warp_sum(wsum);
  const float* xr = logits + row * C;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c] * inv_temp);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(xr[c] * inv_temp - m);
  s = warp_sum(s);
  const float lse = m + __logf(s);
  const int tgt = targets[(size_t)b * ldt + t];
  if (lane == 0) rowloss[row] = lse - xr[tgt] * inv_temp;
  if (dr) {
    const float sc = grad_scale_num * inv_temp / (wsum + 1e-12f);
    for (int c = lane; c < C; c += 32) {
      const float p = __expf(xr[c] * inv_temp - lse);
      dr[c] = (p - (c == tgt ? 1.f : 0.f)) * sc;
    }
  }
}

__global__ void __launch_bounds__(256)
tanh_backward_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = y[i];
    dx[i] = dy[i] * (1.f - v * v);
  }
}

__global__ void __launch_bounds__(256)
lstm_cell_pointwise_bwd_kernel(const float* __restrict__ z, const float* __restrict__ bias,
                               const float* __restrict__ wi, const float* __restrict__ wf,
                               const float* __restrict__ wo, const float* __restrict__ c_prev,
                               const float* __restrict__ dh, const float* __restrict__ dc_in, int B, int H,
                               float forget_bias, float cell_clip, float* __restrict__ dz,
                               float* __restrict__ dc_prev) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, u = idx % H;
  const float* zr = z + (size_t)b * 4 * H;
  const float cp = c_prev[idx];
  float zi = zr[u] + (bias ? bias[u] : 0.f), zg = zr[H + u] + (bias ? bias[H + u] : 0.f);
  float zf = zr[2 * H + u] + (bias ? bias[2 * H + u] : 0.f) + forget_bias;
  float zo = zr[3 * H + u] + (bias ? bias[3 * H + u] : 0.f);
  const float pwi = wi ? wi[u] : 0.f, pwf = wf ? wf[u] : 0.f, pwo = wo ? wo[u] : 0.f;
  zi += pwi * cp; zf += pwf * cp;
  const float i = sigmoidf_(zi), f = sigmoidf_(zf), g = tanhf_(zg);
  const float c_raw = f * cp + i * g;
  float c = c_raw;
  bool pass = true;
  if (cell_clip > 0.f) {
    c = fminf(fmaxf(c_raw, -cell_clip), cell_clip);
    pass = (c_raw >= -cell_clip) && (c_raw <= cell_clip);
  }
  zo += pwo * c;
  const float o = sigmoidf_(zo), tc = tanhf_(c);
  const float dhv = dh[idx];
  const float dzo = dhv * tc * o * (1.f - o);
  float dc = (dc_in ? dc_in[idx] : 0.f) + dhv * o * (1.f - tc * tc) + dzo * pwo;
  if (!pass) dc = 0.f;
  const float dzi = dc * g * i * (1.f - i);
  const float dzg = dc * i * (1.f - g * g);
  const float dzf = dc * cp * f * (1.f - f);
  float* dr = dz + (size_t)b * 4 * H;
  dr[u] = dzi; dr[H + u] = dzg; dr[2 * H + u] = dzf; dr[3 * H + u] = dzo;
  dc_prev[idx] = dc * f + dzi * pwi + dzf * pwf;
}

// rows = steps*B.  dwi[u] += sum_r dz_i[r,u]*c_prev[r,u]; dwf likewise; dwo += dz_o*c[r,u]
// c_all holds steps+1 blocks of [B,H]: block s = state before step s (block 0 = initial state)
__global__ void __launch_bounds__(256)
decoder_peephole_grad_kernel(const float* __restrict__ dz, const float* __restrict__ c_all,
                             int64_t rows, int B, int H, float* dwi, float* dwf, float* dwo) {
  __shared__ float sh[3][8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int u = blockIdx.x * 32 + lane;
  float si = 0.f, sf = 0.f, so = 0.f;
  if (u < H)
    for (int64_t r = (int64_t)blockIdx.y * 8 + wy; r < rows; r += (int64_t)gridDim.y * 8) {
      const float* d = dz + r * 4 * H;
      const float cp = c_all[r * H + u], c = c_all[(r + B) * H + u];
      si = fmaf(d[u], cp, si); sf = fmaf(d[2 * H + u], cp, sf); so = fmaf(d[3 * H + u], c, so);
    }
  sh[0][wy][lane] = si; sh[1][wy][lane] = sf; sh[2][wy][lane] = so;
  __syncthreads();
  if (wy == 0 && u < H) {
#pragma unroll
    for (int j = 1; j < 8; ++j) { si += sh[0][j][lane]; sf += sh[1][j][lane]; so += sh[2][j][lane]; }
    atomicAdd(dwi + u, si); atomicAdd(dwf + u, sf); atomicAdd(dwo + u, so);
  }
}

// dW[ids[r], :] += dx[r, :D]   (dx rows have leading dimension ldx)
__global__ void __launch_bounds__(256)
embedding_grad_kernel(const float* __restrict__ dx, int ldx, const int* __restrict__ ids,
                      int64_t rows, int D, int V, float* __restrict__ dW) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const int id = ids[r];
  if (id < 0 || id >= V) return;
  for (int i = lane; i < D; i += 32) {
    const float v = dx[r * ldx + i];
    if (v != 0.f) atomicAdd(dW + (size_t)id * D + i, v);
  }
}

// tf.nn.dropout on a row-strided matrix with an explicit element numbering: element (r, c) is kept iff the
// counter hash of (seed, idx_base + r*idx_row_stride + c) says so.  add != 0: y += mask*x/keep (backward of a
// dropped operand whose gradient accumulates), else y = mask*x/keep (x == y allowed).
__global__ void __launch_bounds__(256)
dropout_rows_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int64_t rows, int cols,
                    float keep, unsigned long long seed, unsigned long long idx_base,
                    unsigned long long idx_row_stride, int add) {
  const int64_t n = rows * cols;
  const float sc = 1.f / keep;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols; const int c = (int)(i % cols);
    const float v = dropout_keep(seed, idx_base + (unsigned long long)r * idx_row_stride + c, keep) ? x[r * ldx + c] * sc : 0.f;
    float* o = y + r * ldy + c;
    *o = add ? *o + v : v;
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_dropout_rows(const float* x, int ldx, float* y, int ldy, int64_t rows, int cols, float keep_prob,
                               uint64_t seed, uint64_t idx_base, uint64_t idx_row_stride, int add,
                               b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(x && y && rows > 0 && cols > 0 && ldx >= cols && ldy >= cols && keep_prob > 0.f && keep_prob <= 1.f,
               "b2_dropout_rows: bad argument");
  int64_t blocks = (rows * cols + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  dropout_rows_kernel<<<(int)blocks, 256, 0, stream>>>(x, ldx, y, ldy, rows, cols, keep_prob, seed, idx_base,
                                                      idx_row_stride, add);
  B2_LAUNCH_CHECK();
  return B2_OK;
}


extern "C" int b2_sequence_loss(const float* logits, const int32_t* targets, int targets_ld,
                                const int32_t* lengths, int B, int L, int C, float temperature,
                                float grad_scale, float* rowloss, float* dlogits,
                                b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(logits && targets && lengths && rowloss, "b2_sequence_loss: null pointer");
  B2_CHECK_ARG(B > 0 && L > 0 && C > 0 && targets_ld >= L && temperature > 0.f, "b2_sequence_loss: bad shape");
  sequence_loss_kernel<<<cdiv((int64_t)B * L, 8), 256, 0, stream>>>(
      logits, targets, targets_ld, lengths, B, L, C, 1.f / temperature, grad_scale, rowloss, dlogits);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_tanh_backward(const float* dy, const float* y, float* dx, int64_t n, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(dy && y && dx && n > 0, "b2_tanh_backward: bad argument");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  tanh_backward_kernel<<<blocks, 256, 0, stream>>>(dy, y, dx, n);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_lstm_cell_pointwise_backward(const float* z, const float* bias, const float* w_i_diag,
                                               const float* w_f_diag, const float* w_o_diag,
                                               const float* c_prev, const float* dh, const float* dc_in,
                                               int B, int H, float forget_bias, float cell_clip,
                                               float* dz, float* dc_prev, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(z && c_prev && dh && dz && dc_prev && B > 0 && H > 0, "b2_lstm_cell_pointwise_backward: bad argument");
  B2_CHECK_ARG((!w_i_diag && !w_f_diag && !w_o_diag) || (w_i_diag && w_f_diag && w_o_diag),
               "b2_lstm_cell_pointwise_backward: give all three peephole vectors or none");
  lstm_cell_pointwise_bwd_kernel<<<cdiv((int64_t)B * H, 256), 256, 0, stream>>>(
      z, bias, w_i_diag, w_f_diag, w_o_diag, c_prev, dh, dc_in, B, H, forget_bias, cell_clip, dz, dc_prev);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_decoder_peephole_grad(const float* dz, const float* c_all, int steps, int B, int H,
                                        float* dw_i, float* dw_f, float* dw_o, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(dz && c_all && dw_i && dw_f && dw_o && steps > 0 && B > 0 && H > 0,
               "b2_decoder_peephole_grad: bad argument");
  const int64_t rows = (int64_t)steps * B;
  int slabs = (int)cdiv(rows, 64); if (slabs > 128) slabs = 128;
  dim3 grid(cdiv(H, 32), slabs);
  decoder_peephole_grad_kernel<<<grid, 256, 0, stream>>>(dz, c_all, rows, B, H, dw_i, dw_f, dw_o);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_embedding_grad(const float* dx, int ldx, const int32_t* ids, int64_t rows, int D, int V,
                                 float* dW, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(dx && ids && dW && rows > 0 && D > 0 && V > 0 && ldx >= D, "b2_embedding_grad: bad argument");
  embedding_grad_kernel<<<cdiv(rows, 8), 256, 0, stream>>>(dx, ldx, ids, rows, D, V, dW);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
