// Version, error reporting and the host copy of the Philox stream.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace lantern {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

}  // namespace lantern

extern "C" int lantern_version(void) { return (LANTERN_ABI_VERSION << 16) | 1; }

extern "C" const char* lantern_last_error(void) { return lantern::g_err; }

extern "C" void lantern_philox_uniforms(uint64_t seed, uint64_t step, uint32_t item, int32_t n, float* out_host) {
  for (int32_t d = 0; d < n; ++d) out_host[d] = lantern::philox_uniform(seed, step, item, (uint32_t)d);
}
