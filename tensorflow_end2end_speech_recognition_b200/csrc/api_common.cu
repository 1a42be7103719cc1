// Error plumbing + version of the C ABI (include/b2asr.h).
#include "common.cuh"
#include <stdarg.h>

namespace b2 {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launches(int n) { g_launches += (unsigned long long)n; }
unsigned long long get_launches() { return g_launches; }
}  // namespace b2

extern "C" unsigned long long b2_launch_count(void) { return b2::get_launches(); }
extern "C" int b2_version(void) { return 100; }
extern "C" const char* b2_last_error(void) { return b2::g_err; }
extern "C" int b2_device_is_sm100(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

// Beam search lives in beam.cu once built; until then the symbols exist and fail loudly.
#ifndef B2_HAVE_BEAM
extern "C" size_t b2_ctc_beam_workspace_bytes(int, int, int, int) { return 0; }
extern "C" int b2_ctc_beam_decode(const float*, const int32_t*, int, int, int, int, int, int32_t*,
                                  int32_t*, float*, void*, size_t, b2_stream_t) {
  b2::set_error("b2_ctc_beam_decode: not built");
  return B2_ERR_UNSUPPORTED;
}
#endif
