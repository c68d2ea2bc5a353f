// tcgen05 with MN-MAJOR operands (the weight-gradient form): D[m, n] = sum_k A[k + shift][m] * B[k][n] with
// both operands stored [K rows][MN columns] -- channel-minor activation / gradient tensors: a row = 32 channels of one
// pixel, K = pixels -- as 32-column blocks of 128-byte rows written by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
// (32-byte chunks ^ (row & 3): the only layout tcgen05 accepts for MN-major tf32, UMMA layout type 1; with the ordinary
// SWIZZLE_128B descriptor the MMA returns zeros, tc_mn_probe.cu).  Checks the descriptor encoding (a_major / b_major = 1
// in the instruction descriptor; LBO = byte distance of the 32-column MN blocks, SBO = 512 = four K rows) and whether the
// K origin may be advanced by an arbitrary number of rows (the tap shift of the weight gradient).
// usage: tc_gemm_mn
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// MN-major, SWIZZLE_128B: LBO = distance of the 32-element MN blocks, SBO = distance of the 8-row K blocks
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;      // SWIZZLE_128B_BASE32B: the only shared-memory layout of MN-major tf32 operands
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int M = 128, N = 128, KT = 64, AROWS = KT + 16;   // K rows used, rows stored (room for the shift)

__global__ void __launch_bounds__(128, 1)
kern(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C, int shift, int swap_lbo) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* As = reinterpret_cast<float*>(smem);                 // [4 blocks][AROWS][32]
  float* Bs = As + 4 * AROWS * 32;                            // [4 blocks][KT][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + 4 * KT * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect(&bars[0], (uint32_t)(4 * AROWS + 4 * KT) * 128u);
    for (int c = 0; c < 4; ++c) {
      tma2d(As + c * AROWS * 32, &tmA, c * 32, 0, &bars[0]);
      tma2d(Bs + c * KT * 32, &tmB, c * 32, 0, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D = F32, A = B = TF32, a_major = b_major = 1 (MN-major: bits 15, 16), N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t acc = 0;
    for (int k = 0; k < KT / 8; ++k) {
      const uint32_t a0 = smem_u32(As) + (uint32_t)(shift + 8 * k) * 128u;
      const uint32_t b0 = smem_u32(Bs) + (uint32_t)(8 * k) * 128u;
      const uint32_t lba = AROWS * 128u, lbb = KT * 128u;
      const uint64_t da = swap_lbo ? make_desc(a0, 512u, lba) : make_desc(a0, lba, 512u);
      const uint64_t db = swap_lbo ? make_desc(b0, 512u, lbb) : make_desc(b0, lbb, 512u);
      mma_tf32(tmem, da, db, idesc, acc);
      acc = 1;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) C[(size_t)(32 * warp + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(EncodeFn fn, float* base, int cols, int rows) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return tm;
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeFn fn = (EncodeFn)p;
  std::vector<float> A((size_t)AROWS * M), B((size_t)KT * N);
  srand(11);
  auto trunc = [](float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; };
  for (auto& v : A) v = trunc((float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto& v : B) v = trunc((float)rand() / RAND_MAX * 2.f - 1.f);
  float *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap tA = make_map(fn, dA, M, AROWS), tB = make_map(fn, dB, N, KT);
  const size_t smem = (size_t)(4 * AROWS + 4 * KT) * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> Cc((size_t)M * N);
  for (int swap : {0, 1})
    for (int shift : {0, 8, 1, 3, 13}) {
      CK(cudaMemset(dC, 0, Cc.size() * 4));
      kern<<<1, 128, smem>>>(tA, tB, dC, shift, swap);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(Cc.data(), dC, Cc.size() * 4, cudaMemcpyDeviceToHost));
      double worst = 0, rms = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < KT; ++k) ref += (double)A[(size_t)(k + shift) * M + m] * B[(size_t)k * N + n];
          worst = fmax(worst, fabs(ref - Cc[(size_t)m * N + n]));
          rms += ref * ref;
        }
      rms = sqrt(rms / (M * N));
      printf("MN-major A and B, %s, K-row shift %2d: max|err| = %.3e (rms of result %.3f, relative %.2e) %s\n",
             swap ? "LBO = 512 / SBO = block stride" : "LBO = block stride / SBO = 512", shift, worst, rms, worst / rms,
             worst / rms < 1e-5 ? "OK" : "WRONG");
    }
  return 0;
}
