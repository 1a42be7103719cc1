// clock64() rate vs globaltimer: (a) one thread on an idle GPU, (b) the same while every SM runs an FMA loop.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void probe(long long cycles, double* out, int busy) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long g0 = gtimer(); const long long c0 = clock64();
    while (clock64() - c0 < cycles) {}
    const unsigned long long g1 = gtimer(); const long long c1 = clock64();
    out[0] = (double)(c1 - c0) / (double)(g1 - g0);
  } else if (busy) {
    float a = threadIdx.x, b = 1.0001f;
    const long long c0 = clock64();
    while (clock64() - c0 < cycles) {
#pragma unroll
      for (int i = 0; i < 64; ++i) a = fmaf(a, b, 0.5f);
    }
    if (a == 123.f) out[1] = a;
  }
}
int main() {
  double* d; cudaMalloc(&d, 16); double h[2];
  for (int busy = 0; busy < 2; ++busy)
    for (int rep = 0; rep < 3; ++rep) {
      probe<<<busy ? 148 * 4 : 1, busy ? 256 : 32>>>(40000000LL, d, busy);
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("busy=%d clock64 rate %.4f GHz\n", busy, h[0]);
    }
  return 0;
}
