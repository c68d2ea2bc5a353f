// Micro-benchmark: how fast can ONE streaming pass (read [+ write]) over a buffer much larger than L2 go on
// sm_100a when it is fed by (a) LDG.128 from registers with U independent loads in flight per thread, and
// (b) cp.async.bulk (the TMA's 1-D bulk copy) into a shared-memory ring consumed with LDS.128?
// It pins the plateau the LSU-fed elementwise / gather kernels of this library sit on (criterions, sampler)
// against the TMA-fed cost volume, and tells what an elementwise criterion kernel should be built like.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/stream_feed tools/ubench/stream_feed.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// ---- (a) LDG: grid-stride, U float4 loads issued before the first use ----
template <int U, bool WRITE>
__global__ void __launch_bounds__(256) ldg_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n4, float* sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(in + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (WRITE) __stcs(out + i + u * stride, make_float4(v[u].x + 1.f, v[u].y, v[u].z, v[u].w));
      else acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  for (; i < n4; i += stride) {
    const float4 v = __ldcs(in + i);
    if (WRITE) __stcs(out + i, make_float4(v.x + 1.f, v.y, v.z, v.w));
    else acc += v.x + v.y + v.z + v.w;
  }
  if (!WRITE && acc == 123.456f) *sink = acc;
}

// ---- (b) bulk copy into a shared-memory ring ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// One producer thread (the last warp) keeps STAGES chunks of CHUNK bytes in flight; 256 consumer threads read
// each chunk with LDS.128.  WRITE: consumers modify the chunk in place and one thread bulk-stores it.
template <int STAGES, int CHUNK, bool WRITE>
__global__ void __launch_bounds__(288) bulk_kernel(const char* __restrict__ in, char* __restrict__ out, size_t nchunks, float* sink) {
  extern __shared__ __align__(128) char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t first = blockIdx.x, step = gridDim.x;
  if (threadIdx.x >= 256) {
    if (threadIdx.x == 256) {
      int it = 0;
      for (size_t c = first; c < nchunks; c += step, ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) {
          mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
          if (WRITE) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store of this slot has read it
        }
        mbar_expect(&full[s], CHUNK);
        bulk_load(smem + (size_t)s * CHUNK, in + c * CHUNK, CHUNK, &full[s]);
      }
      if (WRITE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    return;
  }
  float acc = 0.f;
  int it = 0;
  for (size_t c = first; c < nchunks; c += step, ++it) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    float4* p = reinterpret_cast<float4*>(smem + (size_t)s * CHUNK);
#pragma unroll
    for (int k = 0; k < CHUNK / 16 / 256; ++k) {
      float4 v = p[threadIdx.x + 256 * k];
      if (WRITE) { v.x += 1.f; p[threadIdx.x + 256 * k] = v; }
      else acc += v.x + v.y + v.z + v.w;
    }
    if (WRITE) {
      // consumers' generic-proxy writes must be visible to the async proxy before the bulk store reads them
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        bulk_store(out + c * CHUNK, p, CHUNK);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    mbar_arrive(&empty[s]);
  }
  if (WRITE && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (!WRITE && acc == 123.456f) *sink = acc;
}

template <class F>
float time_ms(F f, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t bytes = (size_t)1 << 30;   // 1 GiB in, 1 GiB out: far beyond the 126 MB L2
  char *in, *out; float* sink;
  CK(cudaMalloc(&in, bytes)); CK(cudaMalloc(&out, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(in, 1, bytes)); CK(cudaMemset(out, 0, bytes));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t n4 = bytes / 16;
  printf("SMs %d, buffer %zu MiB\n", sms, bytes >> 20);
  printf("cudaMemcpy D2D:            %7.0f GB/s (read+write)\n", 2.0 * bytes / time_ms([&] { cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice); }) / 1e6);
#define RUN_LDG(U, BPS) { \
    float r = time_ms([&] { ldg_kernel<U, false><<<sms * BPS, 256>>>((const float4*)in, (float4*)out, n4, sink); }); \
    float w = time_ms([&] { ldg_kernel<U, true><<<sms * BPS, 256>>>((const float4*)in, (float4*)out, n4, sink); }); \
    printf("LDG.128 U=%d blocks/SM=%d:   read %7.0f GB/s | copy %7.0f GB/s (read+write)\n", U, BPS, bytes / r / 1e6, 2.0 * bytes / w / 1e6); }
  RUN_LDG(1, 8) RUN_LDG(2, 8) RUN_LDG(4, 8) RUN_LDG(8, 8) RUN_LDG(4, 4) RUN_LDG(8, 4) RUN_LDG(8, 2) RUN_LDG(16, 2)
  CK(cudaGetLastError());
#define RUN_BULK(ST, CH, BPS) { \
    const int smem = ST * CH + 2 * ST * 8 + 128; \
    cudaFuncSetAttribute(bulk_kernel<ST, CH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    cudaFuncSetAttribute(bulk_kernel<ST, CH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    float r = time_ms([&] { bulk_kernel<ST, CH, false><<<sms * BPS, 288, smem>>>(in, out, bytes / CH, sink); }); \
    float w = time_ms([&] { bulk_kernel<ST, CH, true><<<sms * BPS, 288, smem>>>(in, out, bytes / CH, sink); }); \
    printf("bulk %d x %5d B, CTAs/SM=%d: read %7.0f GB/s | copy %7.0f GB/s (read+write)\n", ST, CH, BPS, bytes / r / 1e6, 2.0 * bytes / w / 1e6); }
  RUN_BULK(2, 16384, 1) RUN_BULK(4, 16384, 1) RUN_BULK(8, 16384, 1) RUN_BULK(4, 8192, 2) RUN_BULK(4, 16384, 2) RUN_BULK(6, 16384, 2) RUN_BULK(3, 32768, 2)
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  return 0;
}
