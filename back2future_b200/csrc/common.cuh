// Shared host/device helpers for libb2f_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/b2f.h"

namespace b2f {

// ---- thread-local error text, launch counter, debug switches (api.cu) -----------------
void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);          // sets the message, returns status
int cuda_fail(cudaError_t e, const char* where);     // positive cudaError_t + message
void count_launch(int n = 1);
int release_scratch_for_thread();   // criterions.cu
int reserve_scratch_for_thread(size_t bytes);
int costvol_path();   // 0 auto, 1 generic, 2/3: tiled (32-column fwd tiles), 4: 16-column tiles, 5: channel split,
                      // 6/7: force / forbid the software-pipelined forward

#define B2F_CUDA_TRY(expr)                                              \
  do {                                                                  \
    cudaError_t e__ = (expr);                                           \
    if (e__ != cudaSuccess) return ::b2f::cuda_fail(e__, #expr);        \
  } while (0)

// After a kernel launch: catch launch-configuration errors without synchronising.
#define B2F_CHECK_LAUNCH(name)                                          \
  do {                                                                  \
    cudaError_t e__ = cudaGetLastError();                               \
    if (e__ != cudaSuccess) return ::b2f::cuda_fail(e__, name);         \
    ::b2f::count_launch();                                              \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

inline int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) cached_sms = n;
    cached_dev = dev;
  }
  return cached_sms;
}

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Penalty functions of criterions/penalty/*.lua.  eps2 is eps^2 of the Lorentzian.  Error budget against the 1e-4
// parity bar (CUDA C Programming Guide, "Intrinsic Functions"):
//   rsqrtf            <= 2 ulp (2.4e-7 relative) over the full range;
//   __fdividef(x, y)  <= 2 ulp for |y| in [2^-126, 2^126] (every denominator here is >= 1e-6 or a pixel count);
//   __expf(x)         = ex2.approx(x * log2(e)): <= 2 + floor(|1.16 x|) ulp, i.e. the error GROWS with |x|.  The edge
//                       weights evaluate exp(-cs * g) with cs = 20 and g = mean_c |dT| <= 4.7 for ColorNormalize'd
//                       images, so x >= -94 and the bound is ~111 ulp = 1.3e-5 relative -- 8x inside the bar, not
//                       500x; below x = -87.3 the result is denormal and flushed to zero (absolute error < 1.2e-38).
//                       tests/test_gpu_parity.py::test_smoothness_weight_extremes checks exactly that range;
//   the Lorentzian's logarithm is NOT approximated: __logf is absolute-error bounded (2^-21.4 on [0.5, 2]) and
//                       log(1 + x^2/(2 eps^2)) sits next to 1 for small residuals, which would break the bar
//                       element-wise in grad_occ; logf(1 + t) in the reference's own operation order instead
//                       (Lorentzian_function.lua:25-26).  No BASELINE configuration selects this penalty.
template <int KIND>
__device__ __forceinline__ float pen_apply(float x, float eps2) {
  if (KIND == B2F_PENALTY_QUADRATIC) return x * x;
  // one MUFU.RSQ instead of the IEEE sqrt sequence: (x^2+eps) * rsqrt(x^2+eps)
  if (KIND == B2F_PENALTY_L1) {
    const float t = x * x + 1e-6f;
    return t * rsqrtf(t);
  }
  return logf(1.f + 0.5f * ((x * x) / eps2));
}
template <int KIND>
__device__ __forceinline__ float pen_der(float x, float eps2) {
  if (KIND == B2F_PENALTY_QUADRATIC) return 2.f * x;
  if (KIND == B2F_PENALTY_L1) return x * rsqrtf(x * x + 1e-6f);
  return __fdividef(2.f * x, x * x + 2.f * eps2);
}

}  // namespace b2f
