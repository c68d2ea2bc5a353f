// probe: which shared-memory element does tcgen05.mma read as A(m, k) when the A descriptor / instruction descriptor say
// MN-major?  B is a K-major identity (known-good encoding, conv_tc.cu): D[m][n] = A(m, k = n) for n < 8 (one K = 8 MMA).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
constexpr int M = 128, N = 128, ROWS = 64;
// smem A: 4 blocks x [ROWS][32 floats], written by plain stores with the 128-byte swizzle (chunk ^= row & 7)
__global__ void __launch_bounds__(128, 1) probe(float* C, int what, uint32_t abit, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* As = reinterpret_cast<float*>(smem);        // 4 * ROWS * 32
  float* Bs = As + 4 * ROWS * 32;                    // K-major identity: [128 rows][32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + 128 * 32);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * ROWS * 32; i += 128) {
    const int blk = i / (ROWS * 32), r = (i / 32) % ROWS, c = i % 32;
    const float v = what == 0 ? (float)(r + 1) : (float)(blk * 32 + c + 1);
    // layout 2: 16-byte chunks ^ (row & 7) (SWIZZLE_128B); layout 1: 32-byte chunks ^ (row & 3) (SWIZZLE_128B_BASE32B, Swizzle<2,5,2>)
    int cc = c;
    if (layout == 2) cc = (((c >> 2) ^ (r & 7)) << 2) | (c & 3);
    if (layout == 1) cc = (((c >> 3) ^ (r & 3)) << 3) | (c & 7);
    As[(blk * ROWS + r) * 32 + cc] = v;
  }
  for (int i = threadIdx.x; i < 128 * 32; i += 128) {
    const int r = i / 32, c = i % 32;
    const int sw = (c >> 2) ^ (r & 7);
    Bs[r * 32 + sw * 4 + (c & 3)] = (r == c && c < 8) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | abit | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint64_t da = (uint64_t)((smem_u32(As) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
                  ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
    uint64_t db = (uint64_t)((smem_u32(Bs) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    mma_tf32(tmem, da, db, idesc, 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  __syncwarp();
  { uint32_t done; do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)) : "memory"); } while (!done); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tmem + ((uint32_t)(32 * warp) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 8; ++j) C[(32 * warp + lane) * 8 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N) : "memory");
}
int main() {
  float* dC; CK(cudaMalloc(&dC, 128 * 8 * 4));
  const size_t smem = (4 * ROWS * 32 + 128 * 32) * 4 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> R(128 * 8), Cc(128 * 8);
  struct Cfg { uint32_t abit, lbo, sbo, layout; const char* name; };
  const Cfg cfgs[] = {{0, 16, 1024, 2, "K-major reference (no extra bit)"},
                      {1u << 15, ROWS * 128, 512, 1, "a_major=MN, layout 1 (128B_BASE32B), LBO=block, SBO=512"},
                      {1u << 15, 512, ROWS * 128, 1, "a_major=MN, layout 1, LBO=512, SBO=block"},
                      {1u << 15, ROWS * 128, 1024, 1, "a_major=MN, layout 1, LBO=block, SBO=1024"},
                      {1u << 15, 1024, ROWS * 128, 1, "a_major=MN, layout 1, LBO=1024, SBO=block"}};
  for (const Cfg& c : cfgs) {
    probe<<<1, 128, smem>>>(dC, 0, c.abit, c.lbo, c.sbo, c.layout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(R.data(), dC, R.size() * 4, cudaMemcpyDeviceToHost));
    probe<<<1, 128, smem>>>(dC, 1, c.abit, c.lbo, c.sbo, c.layout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(Cc.data(), dC, Cc.size() * 4, cudaMemcpyDeviceToHost));
    printf("%s\n  A(m,k) read from (row,col): ", c.name);
    for (int m : {0, 1, 5, 31, 32, 33, 64, 127}) { printf(" m=%d:", m); for (int k : {0, 1, 7}) printf("(%g,%g)", R[m * 8 + k], Cc[m * 8 + k]); }
    printf("\n");
  }
  return 0;
}
