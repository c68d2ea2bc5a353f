// How fast can TMA move halo'd 2-D/4-D boxes from L2/HBM into shared memory, as a function of the box row
// length?  Every CTA streams boxes {w, rows, ch} of an fp32 (B, C, H, W) tensor through a ring (no consumer
// work), boxes start 16 B before a 128-byte line like the cost-volume halos do.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/tma_feed tools/ubench/tma_feed.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t a, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t ph) {
  uint32_t d;
  do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(d) : "r"(a), "r"(ph) : "memory"); } while (!d);
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
               ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// one thread per CTA issues; NS boxes in flight; each box is waited for (by the same thread) before its slot is reused
__global__ void feed(const __grid_constant__ CUtensorMap tm, int box_bytes, int ns, int nbox, int ntx, int nty, int nch, int tw, int th, int ck, int xoff) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = s32(smem);
  const uint32_t bars = base + ns * box_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ns; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int k = 0; k < nbox + ns; ++k) {
      const int s = k % ns;
      if (k >= ns) mbar_wait(bars + 8 * s, ((k / ns) - 1) & 1);
      if (k < nbox) {
        int t = blockIdx.x + k * gridDim.x;      // box index, raster order
        const int tx = t % ntx; t /= ntx;
        const int ty = t % nty; t /= nty;
        const int c = t % nch; const int b = t / nch;
        mbar_expect(bars + 8 * s, box_bytes);
        tma4(base + s * box_bytes, &tm, tx * tw + xoff, ty * th - 4, c * ck, b, bars + 8 * s);
      }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int B = 8, C = 32, H = 112, W = 256;   // level 3 of the benchmark
  float* d; CK(cudaMalloc(&d, (size_t)B * C * H * W * 4)); CK(cudaMemset(d, 0, (size_t)B * C * H * W * 4));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(feed, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Cfg { int tw, bw, th, rows, ck, xoff; const char* name; };
  const Cfg cfgs[] = {
      {32, 44, 8, 16, 8, -4, "halo 44 x 16 x 8ch (cost-volume X/frame box)"},
      {32, 36, 8, 8, 8, 0, "36 x 8 x 8ch aligned (ref box)"},
      {32, 32, 8, 16, 8, 0, "32 x 16 x 8ch aligned, no halo"},
      {64, 76, 8, 16, 8, -4, "halo 76 x 16 x 8ch (64-column tiles)"},
      {64, 64, 8, 16, 8, 0, "64 x 16 x 8ch aligned"},
      {128, 140, 8, 16, 4, -4, "halo 140 x 16 x 4ch"},
      {256, 256, 8, 16, 2, 0, "256 x 16 x 2ch (full rows)"},
      {32, 44, 8, 8, 1, -4, "44 x 8 x 1ch (one gradOut box, 1.4 KB)"},
  };
  for (const Cfg& c : cfgs) {
    cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)c.bw, (cuuint32_t)c.rows, (cuuint32_t)c.ck, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const int box_bytes = ((c.bw * c.rows * c.ck * 4 + 127) / 128) * 128;
    const int ntx = W / c.tw, nty = H / c.th, nch = C / c.ck;
    const int total = ntx * nty * nch * B;
    for (int per_sm = 1; per_sm <= 2; ++per_sm) {
      int ns = (170 * 1024 / per_sm) / box_bytes; if (ns > 16) ns = 16; if (ns < 1) ns = 1;
      const int grid = sms * per_sm;
      const int nbox = total / grid;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int smem = ns * box_bytes + 8 * ns + 64;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) feed<<<grid, 32, smem>>>(tm, box_bytes, ns, nbox, ntx, nty, nch, c.tw, c.th, c.ck, c.xoff);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = 5.0 * (double)nbox * grid * c.bw * c.rows * c.ck * 4;
      printf("%-48s CTAs/SM %d ring %2d x %6d B: %7.1f us/launch  %6.2f TB/s into smem\n", c.name, per_sm, ns, box_bytes, ms * 1e3 / 5, bytes / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
