// Library-level entry points of libb2f_cuda.so: version, error text, counters, test hooks.
#include "common.cuh"

#include <cstring>

namespace b2f {

namespace {
thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;
thread_local int g_force_generic = 0;
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
  return (int)e;
}

void count_launch(int n) { g_launches += n; }
int costvol_path() { return g_force_generic; }

}  // namespace b2f

extern "C" {

int b2f_abi_version(void) { return B2F_ABI_VERSION; }

const char* b2f_last_error(void) { return b2f::g_err; }

const char* b2f_status_string(int status) {
  switch (status) {
    case B2F_OK: return "ok";
    case B2F_EINVAL: return "invalid argument";
    case B2F_EUNSUPPORTED: return "unsupported configuration";
    case B2F_ENOMEM: return "out of memory";
    case B2F_EALIGN: return "misaligned pointer";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "unknown status";
}

int b2f_release_scratch(void) { return b2f::release_scratch_for_thread(); }

int b2f_debug_costvol_path(int mode) {
  int prev = b2f::g_force_generic;
  b2f::g_force_generic = mode;
  return prev;
}

int b2f_zero_async(void* ptr, size_t bytes, b2f_stream_t stream) {
  if (!ptr && bytes) return b2f::fail(B2F_EINVAL, "zero_async: NULL pointer");
  if (!bytes) return B2F_OK;
  cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return b2f::cuda_fail(e, "cudaMemsetAsync");
  return B2F_OK;
}

int64_t b2f_launch_count(int reset) {
  int64_t v = b2f::g_launches;
  if (reset) b2f::g_launches = 0;
  return v;
}

}  // extern "C"
