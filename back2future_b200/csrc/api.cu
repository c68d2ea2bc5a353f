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
int b2f_reserve_scratch(size_t bytes) { return b2f::reserve_scratch_for_thread(bytes); }

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

int b2f_copy2d_async(float* dst, int64_t dst_row_stride, const float* src, int64_t src_row_stride, int64_t row_elems,
                     int64_t rows, b2f_stream_t stream) {
  if (rows < 0 || row_elems < 0) return b2f::fail(B2F_EINVAL, "copy2d_async: negative extent");
  if (!rows || !row_elems) return B2F_OK;
  if (!dst || !src) return b2f::fail(B2F_EINVAL, "copy2d_async: NULL pointer");
  if (dst_row_stride < row_elems || src_row_stride < row_elems)
    return b2f::fail(B2F_EINVAL, "copy2d_async: row stride smaller than the row");
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dst_row_stride * 4, src, (size_t)src_row_stride * 4, (size_t)row_elems * 4,
                                    (size_t)rows, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return b2f::cuda_fail(e, "cudaMemcpy2DAsync");
  return B2F_OK;
}

// ---- .flo files: host-side wire format of flow fields (flowExtensions.lua:254-287) ----
namespace {
constexpr float kFloTag = 202021.25f;
struct File {
  FILE* f;
  explicit File(const char* path, const char* mode) : f(path ? fopen(path, mode) : nullptr) {}
  ~File() { if (f) fclose(f); }
};
int flo_header(FILE* f, const char* path, int* w, int* h) {
  float tag = 0.f;
  int32_t dims[2] = {0, 0};
  if (fread(&tag, 4, 1, f) != 1 || fread(dims, 4, 2, f) != 2) return b2f::fail(B2F_EINVAL, "flo: %s: truncated header", path);
  if (tag != kFloTag) return b2f::fail(B2F_EINVAL, "flo: unable to read %s perhaps bigendian error", path);
  if (dims[0] <= 0 || dims[1] <= 0 || dims[0] > 99999 || dims[1] > 99999)
    return b2f::fail(B2F_EINVAL, "flo: %s: bad size %d x %d", path, dims[0], dims[1]);
  *w = dims[0];
  *h = dims[1];
  return B2F_OK;
}
}  // namespace

int b2f_flo_write(const char* path, const float* flow_chw, int h, int w) {
  if (!path || !flow_chw || h <= 0 || w <= 0) return b2f::fail(B2F_EINVAL, "flo_write: bad argument");
  File fh(path, "wb");
  if (!fh.f) return b2f::fail(B2F_EINVAL, "flo_write: cannot open %s", path);
  const int32_t dims[2] = {w, h};
  if (fwrite(&kFloTag, 4, 1, fh.f) != 1 || fwrite(dims, 4, 2, fh.f) != 2) return b2f::fail(B2F_EINVAL, "flo_write: %s: write failed", path);
  const size_t hw = (size_t)h * w;
  float row[2 * 1024];
  for (size_t i = 0; i < hw;) {   // (2,h,w) -> interleaved (u,v), F:permute(2,3,1) in the reference
    const size_t n = hw - i < 1024 ? hw - i : 1024;
    for (size_t j = 0; j < n; ++j) {
      row[2 * j] = flow_chw[i + j];
      row[2 * j + 1] = flow_chw[hw + i + j];
    }
    if (fwrite(row, 8, n, fh.f) != n) return b2f::fail(B2F_EINVAL, "flo_write: %s: write failed", path);
    i += n;
  }
  return B2F_OK;
}

int b2f_flo_read_header(const char* path, int* w, int* h) {
  if (!path || !w || !h) return b2f::fail(B2F_EINVAL, "flo_read_header: bad argument");
  File fh(path, "rb");
  if (!fh.f) return b2f::fail(B2F_EINVAL, "flo: cannot open %s", path);
  return flo_header(fh.f, path, w, h);
}

int b2f_flo_read(const char* path, float* flow_chw, int h, int w) {
  if (!path || !flow_chw) return b2f::fail(B2F_EINVAL, "flo_read: bad argument");
  File fh(path, "rb");
  if (!fh.f) return b2f::fail(B2F_EINVAL, "flo: cannot open %s", path);
  int fw = 0, fhh = 0;
  int rc = flo_header(fh.f, path, &fw, &fhh);
  if (rc) return rc;
  if (fw != w || fhh != h) return b2f::fail(B2F_EINVAL, "flo_read: %s is %d x %d, buffer is %d x %d", path, fw, fhh, w, h);
  const size_t hw = (size_t)h * w;
  float row[2 * 1024];
  for (size_t i = 0; i < hw;) {
    const size_t n = hw - i < 1024 ? hw - i : 1024;
    if (fread(row, 8, n, fh.f) != n) return b2f::fail(B2F_EINVAL, "flo_read: %s: truncated data", path);
    for (size_t j = 0; j < n; ++j) {
      flow_chw[i + j] = row[2 * j];
      flow_chw[hw + i + j] = row[2 * j + 1];
    }
    i += n;
  }
  return B2F_OK;
}

int64_t b2f_launch_count(int reset) {
  int64_t v = b2f::g_launches;
  if (reset) b2f::g_launches = 0;
  return v;
}

}  // extern "C"
