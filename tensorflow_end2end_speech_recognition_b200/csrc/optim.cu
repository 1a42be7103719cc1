// Per-tensor gradient clipping, TF-1.x optimizer update rules and small
// data-movement helpers of the training step (all HBM-bound elementwise work).
//
// b2_clip_by_norm_multi   <- tf.clip_by_norm per gradient tensor, models/model_base.py:148-152
// b2_optimizer_step_multi <- tf.train.{GradientDescent,Momentum,Adagrad,Adadelta,Adam,RMSProp}
//                            Optimizer as selected by models/model_base.py:12-20,68-95
// b2_transpose_01         <- tf.transpose(inputs,[1,0,2]) at models/encoders/core/blstm.py:279
// b2_colsum               <- bias gradients (sum over T*B rows)
#include "common.cuh"

namespace b2 {

__global__ void __launch_bounds__(256)
sumsq_multi_kernel(float* const* __restrict__ grads, const int64_t* __restrict__ sizes,
                   float* __restrict__ norms) {
  const int k = blockIdx.y;
  const float* g = grads[k];
  const int64_t n = sizes[k];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = sh[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(&norms[k], s);
  }
}

__global__ void __launch_bounds__(256)
clip_scale_multi_kernel(float* const* __restrict__ grads, const int64_t* __restrict__ sizes,
                        const float* __restrict__ norms, float clip_norm, float post_scale) {
  const int k = blockIdx.y;
  float* g = grads[k];
  const int64_t n = sizes[k];
  float scale = post_scale;
  if (clip_norm > 0.f) {
    const float nrm = sqrtf(norms[k]);
    scale *= clip_norm / fmaxf(nrm, clip_norm);    // tf.clip_by_norm
  }
  if (scale == 1.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    g[i] *= scale;
}

__global__ void __launch_bounds__(256)
optimizer_multi_kernel(int kind, float* const* __restrict__ params, float* const* __restrict__ grads,
                       float* const* __restrict__ state0, float* const* __restrict__ state1,
                       const int64_t* __restrict__ sizes, float lr, float adam_lr_t) {
  const int k = blockIdx.y;
  float* w = params[k];
  const float* g = grads[k];
  float* s0 = state0 ? state0[k] : nullptr;
  float* s1 = state1 ? state1[k] : nullptr;
  const int64_t n = sizes[k];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float wi = w[i];
    switch (kind) {
      case B2_OPT_SGD: wi -= lr * gi; break;
      case B2_OPT_MOMENTUM: { const float a = 0.9f * s0[i] + gi; s0[i] = a; wi -= lr * a; } break;
      case B2_OPT_NESTEROV: { const float a = 0.9f * s0[i] + gi; s0[i] = a; wi -= lr * (gi + 0.9f * a); } break;
      case B2_OPT_ADAGRAD: { const float a = s0[i] + gi * gi; s0[i] = a; wi -= lr * gi * rsqrtf(a); } break;
      case B2_OPT_ADADELTA: {
        const float a = 0.95f * s0[i] + 0.05f * gi * gi;
        const float u = sqrtf(s1[i] + 1e-8f) * rsqrtf(a + 1e-8f) * gi;
        s0[i] = a; s1[i] = 0.95f * s1[i] + 0.05f * u * u; wi -= lr * u;
      } break;
      case B2_OPT_ADAM: {
        const float m = 0.9f * s0[i] + 0.1f * gi;
        const float v = 0.999f * s1[i] + 0.001f * gi * gi;
        s0[i] = m; s1[i] = v; wi -= adam_lr_t * m / (sqrtf(v) + 1e-8f);
      } break;
      case B2_OPT_RMSPROP: {
        const float ms = 0.9f * s0[i] + 0.1f * gi * gi;
        s0[i] = ms; wi -= lr * gi * rsqrtf(ms + 1e-10f);
      } break;
    }
    w[i] = wi;
  }
}

// y_k += alpha * x_k for a list of tensors (weight-decay gradient, models/ctc/ctc.py:280-286)
__global__ void __launch_bounds__(256)
axpy_multi_kernel(float* const* __restrict__ xs, float* const* __restrict__ ys,
                  const int64_t* __restrict__ sizes, float alpha) {
  const int k = blockIdx.y;
  const float* x = xs[k];
  float* y = ys[k];
  const int64_t n = sizes[k];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

// [d0, d1, d2] -> [d1, d0, d2]
__global__ void __launch_bounds__(256)
transpose01_kernel(const float* __restrict__ x, float* __restrict__ y, int d0, int d1, int d2) {
  const int64_t n = (int64_t)d0 * d1 * d2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d2);
    const int64_t r = i / d2;
    const int b = (int)(r % d0);      // output index i = (a*d0 + b)*d2 + c with a in d1
    const int a = (int)(r / d0);
    y[i] = x[((int64_t)b * d1 + a) * d2 + c];
  }
}

// out[n] (+)= sum_m X[m, n]; grid.x over column tiles of 32, grid.y row slabs
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int64_t M, int N, int ldx, float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (col < N)
    for (int64_t m = (int64_t)blockIdx.y * 8 + wy; m < M; m += (int64_t)gridDim.y * 8)
      s += X[m * ldx + col];
  sh[wy][lane] = s;
  __syncthreads();
  if (wy == 0) {
#pragma unroll
    for (int j = 1; j < 8; ++j) s += sh[j][lane];
    if (col < N) atomicAdd(&out[col], s);
  }
}

static int grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

// Tower mean (utils/training/multi_gpu.py:13-48): dst[i] = (1/n_src) * sum_k srcs[k][i].  dst may alias
// srcs[0].  One streaming pass: n_src reads + 1 write per element, float4 when everything is 16-byte aligned.
constexpr int kMaxTowers = 16;
struct TowerPtrs { const float* p[kMaxTowers]; };

template <bool VEC>
__global__ void __launch_bounds__(256)
tower_mean_kernel(TowerPtrs srcs, int n_src, float* __restrict__ dst, int64_t n, float inv) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 a = ((const float4*)srcs.p[0])[i];
      for (int k = 1; k < n_src; ++k) {
        const float4 b = ((const float4*)srcs.p[k])[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
      ((float4*)dst)[i] = a;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float a = srcs.p[0][i];
      for (int k = 1; k < n_src; ++k) a += srcs.p[k][i];
      dst[i] = a * inv;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float a = srcs.p[0][i];
      for (int k = 1; k < n_src; ++k) a += srcs.p[k][i];
      dst[i] = a * inv;
    }
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_tower_mean(const float* const* srcs_host, int n_src, float* dst, int64_t n,
                             b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(srcs_host && dst && n > 0, "b2_tower_mean: bad argument");
  B2_CHECK_ARG(n_src >= 1 && n_src <= kMaxTowers, "b2_tower_mean: 1..%d towers, got %d", kMaxTowers, n_src);
  TowerPtrs tp;
  bool vec = ((uintptr_t)dst & 15) == 0;
  for (int k = 0; k < n_src; ++k) {
    B2_CHECK_ARG(srcs_host[k] != nullptr, "b2_tower_mean: tower %d is null", k);
    tp.p[k] = srcs_host[k];
    vec = vec && (((uintptr_t)srcs_host[k] & 15) == 0);
  }
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  if (vec) tower_mean_kernel<true><<<(int)blocks, 256, 0, stream>>>(tp, n_src, dst, n, 1.f / n_src);
  else tower_mean_kernel<false><<<(int)blocks, 256, 0, stream>>>(tp, n_src, dst, n, 1.f / n_src);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_clip_by_norm_multi(float* const* grads, const int64_t* sizes, int n,
                                     float clip_norm, float post_scale, float* norms,
                                     b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(grads && sizes && norms && n > 0, "b2_clip_by_norm_multi: bad argument");
  B2_CUDA(cudaMemsetAsync(norms, 0, (size_t)n * sizeof(float), stream));
  dim3 grid(64, n);
  if (clip_norm > 0.f) {
    sumsq_multi_kernel<<<grid, 256, 0, stream>>>(grads, sizes, norms);
    B2_LAUNCH_CHECK();
  }
  clip_scale_multi_kernel<<<grid, 256, 0, stream>>>(grads, sizes, norms, clip_norm, post_scale);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_axpy_multi(float* const* xs, float* const* ys, const int64_t* sizes, int n,
                            float alpha, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(xs && ys && sizes && n > 0, "b2_axpy_multi: bad argument");
  dim3 grid(64, n);
  axpy_multi_kernel<<<grid, 256, 0, stream>>>(xs, ys, sizes, alpha);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_optimizer_step_multi(int kind, float* const* params, float* const* grads,
                                       float* const* state0, float* const* state1,
                                       const int64_t* sizes, int n, float learning_rate,
                                       int64_t step, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(params && grads && sizes && n > 0, "b2_optimizer_step_multi: bad argument");
  B2_CHECK_ARG(kind >= B2_OPT_SGD && kind <= B2_OPT_RMSPROP, "unknown optimizer kind %d", kind);
  B2_CHECK_ARG(kind == B2_OPT_SGD || state0, "optimizer kind %d needs state0", kind);
  B2_CHECK_ARG((kind != B2_OPT_ADAM && kind != B2_OPT_ADADELTA) || state1,
               "optimizer kind %d needs state1", kind);
  float lr_t = learning_rate;
  if (kind == B2_OPT_ADAM) {
    const double t = (double)(step < 1 ? 1 : step);
    lr_t = (float)(learning_rate * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t)));
  }
  dim3 grid(148 * 2, n);
  optimizer_multi_kernel<<<grid, 256, 0, stream>>>(kind, params, grads, state0, state1, sizes,
                                                   learning_rate, lr_t);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_transpose_01(const float* x, float* y, int d0, int d1, int d2,
                               b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(x && y && d0 > 0 && d1 > 0 && d2 > 0, "b2_transpose_01: bad argument");
  transpose01_kernel<<<grid_for((int64_t)d0 * d1 * d2), 256, 0, stream>>>(x, y, d0, d1, d2);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_colsum(const float* X, int64_t M, int N, int ldx, float* out, int accumulate,
                         b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(X && out && M > 0 && N > 0, "b2_colsum: bad argument");
  if (!accumulate) B2_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), stream));
  int slabs = (int)((M + 255) / 256);
  if (slabs > 256) slabs = 256;
  dim3 grid(cdiv(N, 32), slabs);
  colsum_kernel<<<grid, 256, 0, stream>>>(X, M, N, ldx, out);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
