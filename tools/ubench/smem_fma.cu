// Micro-benchmarks that pin the sm_100a constants the cost-volume kernels are designed around:
//   * LDS.128 wavefront cost under multicast (several lanes reading the same 16 bytes)
//   * plain FFMA vs packed fma.rn.f32x2 issue rate
//   * scattered vs coalesced STG.128 cost
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/smem_fma tools/ubench/smem_fma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

// pattern -> per-lane byte offset of a 16-byte read
__device__ __forceinline__ int pat_off(int pat, int lane) {
  switch (pat) {
    case 0: return lane * 16;                                  // 32 distinct, conflict free (512 B)
    case 1: return (lane & 7) * 16;                            // 8 distinct, replicated across quarter-warps
    case 2: return (lane >> 2) * 16;                           // 8 distinct, 2 distinct inside each quarter
    case 3: return ((lane & 7) + (lane >> 3)) * 176;           // 11 distinct rows, pitch 11*16 B
    case 4: return 0;                                          // all lanes same address
    case 5: return (lane & 15) * 16;                           // 16 distinct (256 B), replicated across halves
    case 6: return ((lane & 3) + 4 * ((lane >> 2) & 1)) * 16 + ((lane >> 3)) * 176;  // 2 strips x ... generic
    case 7: return (lane & 7) * 176 + (lane >> 3) * 16;        // 8 rows x 4 strips(16B): all distinct, odd pitch
    default: return lane * 16;
  }
}

__global__ void lds128_kernel(int pat, float* out, long long* cyc) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float* p = sm + pat_off(pat, lane) / 4;
  const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  unsigned a0 = 0, a1 = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned x, y, z, w;
      asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w)
                   : "r"(addr + (((unsigned)(i + k) & 1u) << 13)) : "memory");
      a0 ^= x ^ y; a1 ^= z ^ w;
    }
  }
  float4 acc = make_float4(__uint_as_float(a0), __uint_as_float(a1), 0, 0);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MODE>  // 0: FFMA, 1: fma.rn.f32x2
__global__ void fma_kernel(float* out, long long* cyc, float a, float b) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
  float bb[4] = {b, b + 1, b + 2, b + 3};
  float aa[4] = {a, a + 1, a + 2, a + 3};
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(aa[i & 3], bb[(i >> 2) & 3], acc[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        unsigned long long d, x, y, c;
        asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(acc[i]), "f"(acc[i + 1]));
        asm("mov.b64 %0, {%1,%2};" : "=l"(x) : "f"(aa[i & 3]), "f"(aa[(i + 1) & 3]));
        asm("mov.b64 %0, {%1,%2};" : "=l"(y) : "f"(bb[(i >> 2) & 3]), "f"(bb[(i >> 2) & 3]));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(c));
        asm("mov.b64 {%0,%1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(d));
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// STG.128 patterns: 0 = 512 B contiguous per warp; 1 = 32 lanes x 16 B in 32 different 128-B lines (stride 4 KB);
// 2 = 8 lines x (4 x 16 B at stride 32 B); 3 = 32 lines x 32 B done as two instrs (q=0,1)
__global__ void stg_kernel(int pat, float4* buf, long long* cyc, int reps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t wbase = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32 * 1024;  // 32 x 4 KB region per warp (in float4: /16)
  float4 v = make_float4(lane, warp, blockIdx.x, 1);
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    size_t o;
    if (pat == 0) o = (size_t)lane + (size_t)(r & 255) * 32;                    // contiguous 512 B, advancing
    else if (pat == 1) o = (size_t)lane * 256 + (r & 255);                      // 32 lines (4 KB apart)
    else if (pat == 2) o = (size_t)(lane >> 2) * 256 + (lane & 3) * 2 + (r & 1) + (size_t)((r >> 1) & 31) * 8;
    else o = (size_t)lane * 256 + (r & 255);
    buf[wbase / 16 * 16 / 16 + o] = v;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc; long long h[1024];
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 8192));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d\n", sms);
  for (int warps = 4; warps <= 16; warps *= 2)
    for (int pat = 0; pat < 8; ++pat) {
      lds128_kernel<<<sms, warps * 32, 32768>>>(pat, out, cyc);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
      double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
      printf("LDS.128 pat %d warps/SM %2d: %.2f cycles per warp-instr (SM-wide: %.2f cyc/instr)\n", pat, warps, c / ITERS, c / ITERS / warps);
    }
  for (int warps = 4; warps <= 16; warps *= 2) {
    fma_kernel<0><<<sms, warps * 32>>>(out, cyc, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
    double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    printf("FFMA   warps/SM %2d: %.3f FMA/clk/SM\n", warps, (double)ITERS * 32 * 32 * warps / c);
    fma_kernel<1><<<sms, warps * 32>>>(out, cyc, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
    c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    printf("FFMA2  warps/SM %2d: %.3f FMA/clk/SM\n", warps, (double)ITERS * 32 * 32 * warps / c);
  }
  float4* buf; CK(cudaMalloc(&buf, (size_t)sms * 8 * 32 * 4096 * 4));
  for (int pat = 0; pat < 3; ++pat) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 2048;
    stg_kernel<<<sms, 256>>>(pat, buf, cyc, reps);
    cudaEventRecord(e0);
    stg_kernel<<<sms, 256>>>(pat, buf, cyc, reps);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
    double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    printf("STG.128 pat %d: %.2f cycles per warp-instr per SM (8 warps), %.1f GB/s\n", pat, c / reps / 8, (double)sms * 8 * reps * 512 / ms / 1e6);
  }
  return 0;
}
