// tcgen05 data point: a TF32 tensor-core GEMM tile written directly in PTX for sm_100a, in the
// form the 3x3 convolution of the conv trunk would use it --
//   D[128 pixels, N channels] += A[128, K] * B[N, K]^T,  both operands K-major in shared memory with the 128-byte
//   swizzle (TMA-written), accumulator in TMEM, one elected thread issuing tcgen05.mma.cta_group::1.kind::tf32,
//   completion through tcgen05.commit -> mbarrier, epilogue with tcgen05.ld.32x32b.
// Measures (a) correctness of the descriptors against a CPU double GEMM, (b) the error of one TF32 pass versus the
// three-pass split  a = a_hi + a_lo:  a_hi b_hi + a_lo b_hi + a_hi b_lo  (what fp32 parity at 1e-4 through a six-layer
// decoder needs), (c) whether a descriptor may start at a row that is not a multiple of 8 (the +-1 pixel tap shift of an
// implicit GEMM over channel-minor activations: start address + 128 B * shift, base_offset = shift), and (d) the
// sustained MMA rate of a persistent loop over resident tiles.
// usage: tc_gemm [K=256] [N=128] [iters=2000]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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

// K-major, SWIZZLE_128B: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO = 1024 >> 4 = 64), LBO = 1 (ignored),
// version 1 (bits 46-47), layout type 2 (bits 61-63), base offset (bits 49-51)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int M = 128;
// modes: 0 = one TF32 pass on the raw fp32 data; 1 = three-pass split (A_lo, B_lo supplied); shift: A row offset
template <int N>
__global__ void __launch_bounds__(128, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo, float* __restrict__ C,
               int K, int mode, int shift, int iters, int bo_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nk = K / 32;                                   // 32-float (128-byte) K chunks
  const int a_rows = M + 8;                                // room for the row shift
  float* As = reinterpret_cast<float*>(smem);              // [nk][a_rows][32]
  float* Al = As + (size_t)nk * a_rows * 32;
  float* Bs = Al + (size_t)nk * a_rows * 32;               // [nk][N][32]
  float* Bl = Bs + (size_t)nk * N * 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bl + (size_t)nk * N * 32);
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
    const uint32_t bytes = (uint32_t)nk * (a_rows + N) * 128u * (mode ? 2u : 1u);
    mbar_expect(&bars[0], bytes);
    for (int c = 0; c < nk; ++c) {
      tma2d(As + (size_t)c * a_rows * 32, &tmA, c * 32, 0, &bars[0]);
      tma2d(Bs + (size_t)c * N * 32, &tmB, c * 32, 0, &bars[0]);
      if (mode) {
        tma2d(Al + (size_t)c * a_rows * 32, &tmAlo, c * 32, 0, &bars[0]);
        tma2d(Bl + (size_t)c * N * 32, &tmBlo, c * 32, 0, &bars[0]);
      }
    }
    mbar_wait(&bars[0], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      uint32_t acc = 0;
      const int npass = mode ? 3 : 1;
      for (int p = 0; p < npass; ++p) {
        const float* a = (p == 1) ? Al : As;      // passes: hi*hi, lo*hi, hi*lo
        const float* b = (p == 2) ? Bl : Bs;
        for (int c = 0; c < nk; ++c) {
          const uint32_t a0 = smem_u32(a + (size_t)c * a_rows * 32) + (uint32_t)shift * 128u;
          const uint32_t b0 = smem_u32(b + (size_t)c * N * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k) {            // 8 TF32 = 32 bytes per MMA
            mma_tf32(tmem, make_desc(a0 + 32u * k, bo_mode == 0 ? shift : (bo_mode == 1 ? 0 : (8 - shift) & 7)), make_desc(b0 + 32u * k, 0), idesc, acc);
            acc = 1;
          }
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) = rows of the tile
  if (blockIdx.x == 0) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) C[(size_t)(32 * warp + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn fn, float* base, int K, int rows, int box_rows) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return tm;
}

template <int N>
static void run(int K, int iters) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeFn fn = (EncodeFn)p;
  const int a_rows = M + 8;
  std::vector<float> A((size_t)a_rows * K), B((size_t)N * K), Alo(A.size()), Blo(B.size()), Ahi(A.size()), Bhi(B.size());
  srand(7);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  auto split = [](const std::vector<float>& x, std::vector<float>& hi, std::vector<float>& lo) {
    for (size_t i = 0; i < x.size(); ++i) {
      uint32_t u;
      memcpy(&u, &x[i], 4);
      u &= 0xFFFFE000u;                           // keep 10 mantissa bits: exactly representable in TF32
      memcpy(&hi[i], &u, 4);
      lo[i] = x[i] - hi[i];
    }
  };
  split(A, Ahi, Alo);
  split(B, Bhi, Blo);
  float *dA, *dAh, *dAl, *dB, *dBh, *dBl, *dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dAh, A.size() * 4)); CK(cudaMalloc(&dAl, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dBh, B.size() * 4)); CK(cudaMalloc(&dBl, B.size() * 4));
  CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAh, Ahi.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAl, Alo.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBh, Bhi.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBl, Blo.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(K / 32) * (a_rows + N) * 128 * 2 + 64 + 1024;
  auto kern = tc_gemm_kernel<N>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> Cc((size_t)M * N);
  auto check = [&](const char* what, int shift, bool hi_only) {
    CK(cudaMemcpy(Cc.data(), dC, Cc.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, rms = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)(hi_only ? Ahi : A)[(size_t)(m + shift) * K + k] * (hi_only ? Bhi : B)[(size_t)n * K + k];
        worst = fmax(worst, fabs(ref - Cc[(size_t)m * N + n]));
        rms += ref * ref;
      }
    rms = sqrt(rms / (M * N));
    printf("%-46s K=%d N=%d shift=%d  max|err| = %.3e  (rms of result %.3f, relative %.2e)\n", what, K, N, shift, worst, rms, worst / rms);
  };
  CUtensorMap tA = make_map(fn, dA, K, a_rows, a_rows), tB = make_map(fn, dB, K, N, N);
  CUtensorMap tAh = make_map(fn, dAh, K, a_rows, a_rows), tAl = make_map(fn, dAl, K, a_rows, a_rows);
  CUtensorMap tBh = make_map(fn, dBh, K, N, N), tBl = make_map(fn, dBl, K, N, N);
  for (int bo : {0, 1, 2})
    for (int shift : {0, 1, 2, 5, 8}) {
      CK(cudaMemset(dC, 0, Cc.size() * 4));
      kern<<<1, 128, smem>>>(tAh, tAl, tBh, tBl, dC, K, 0, shift, 1, bo);
      CK(cudaDeviceSynchronize());
      printf("base_offset mode %d (0: = shift, 1: 0, 2: 8 - shift)  ", bo);
      check("one TF32 pass, pre-truncated data", shift, true);
    }
  kern<<<1, 128, smem>>>(tA, tAl, tB, tBl, dC, K, 0, 0, 1, 0);
  CK(cudaDeviceSynchronize());
  check("one TF32 pass on raw fp32 vs fp32 product", 0, false);
  for (int shift : {0, 1}) {
    kern<<<1, 128, smem>>>(tAh, tAl, tBh, tBl, dC, K, 1, shift, 1, 1);
    CK(cudaDeviceSynchronize());
    check("three-pass split vs fp32 product", shift, false);
  }
  // sustained rate: every SM runs `iters` repetitions of the tile's MMAs on resident operands
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int mode : {0, 1}) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<sms, 128, smem>>>(tAh, tAl, tBh, tBl, dC, K, mode, 0, 10, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<<<sms, 128, smem>>>(tAh, tAl, tBh, tBl, dC, K, mode, 0, iters, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * M * N * K * (mode ? 3 : 1) * (double)iters * sms;
    printf("sustained, %d SMs x %d iterations, %s: %.3f ms, %.1f TFLOP/s of TF32 MMA = %.1f TFLOP/s of fp32-equivalent work\n", sms,
           iters, mode ? "three passes" : "one pass", ms, flop / ms / 1e9, flop / (mode ? 3 : 1) / ms / 1e9);
  }
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 64;
  const int N = argc > 2 ? atoi(argv[2]) : 128;
  const int iters = argc > 3 ? atoi(argv[3]) : 2000;
  if (K % 32 || K > 96) { printf("K must be 32, 64 or 96 (operands stay resident in shared memory)\n"); return 1; }
  if (N == 128) run<128>(K, iters);
  else if (N == 64) run<64>(K, iters);
  else if (N == 256) run<256>(K, iters);
  else { printf("N must be 64, 128 or 256\n"); return 1; }
  return 0;
}
