// Shared helpers for the b2asr CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/b2asr.h"

namespace b2 {

void set_error(const char* fmt, ...);
void count_launches(int n);   // bookkeeping for bench.py's gpu_launches (b2_launch_count)

#define B2_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      b2::set_error(__VA_ARGS__);                                 \
      return B2_ERR_INVALID;                                      \
    }                                                             \
  } while (0)

#define B2_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      b2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,             \
                    cudaGetErrorString(e__));                                 \
      return B2_ERR_CUDA;                                                     \
    }                                                                         \
  } while (0)

#define B2_LAUNCH_CHECK() do { b2::count_launches(1); B2_CUDA(cudaGetLastError()); } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// log(exp(a)+exp(b)+exp(c)) with -inf handling (CTC lattice).
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}
// branch-free form (-inf safe): lets the compiler interleave many independent chains
__device__ __forceinline__ float lse3_nb(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, fmaxf(b, c)), -1e30f);
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}
__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + __logf(__expf(a - m) + __expf(b - m));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// tanh via exp; accurate to ~1e-7 abs, saturates cleanly.
__device__ __forceinline__ float tanhf_(float x) {
  float e = __expf(-2.f * fabsf(x));
  float r = (1.f - e) / (1.f + e);
  return copysignf(r, x);
}

// Counter-based dropout mask (DropoutWrapper(output_keep_prob), blstm.py:308-311):
// keep iff hash(seed, idx) < keep_prob.  Same function is restated in numpy by
// the tests, so the oracle can apply the identical mask.
__host__ __device__ __forceinline__ uint32_t mix32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
__host__ __device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, float keep_prob) {
  // 24-bit uniform in [0,1)
  return (float)(mix32(seed, idx) >> 8) * (1.0f / 16777216.0f) < keep_prob;
}

}  // namespace b2
