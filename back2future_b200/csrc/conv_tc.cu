// 3x3 convolution (stride 1, pad 1) + bias + LeakyReLU on the 5th-generation tensor cores (tcgen05 / TMEM) for sm_100a.
// The decoders of models/pwc.lua:76-85 are 93 % of the network's arithmetic: nn.SpatialConvolution
// (nIn, nOut, 3, 3, 1, 1, 1, 1) + nn.LeakyReLU(0.2) as an implicit GEMM
//
//     D[128 pixels, N = Cout] += sum over taps (ky, kx), 32-channel chunks c:   A_tap,c [128, 32] * W_tap,c [N, 32]^T
//
// with kind::tf32 MMAs.  One TF32 pass loses 13 mantissa bits (measured 2.8e-3 relative on a 64-term dot product,
// tools/ubench/tc_gemm.cu), far outside the 1e-4 parity bar through a six-layer decoder, so every operand is SPLIT:
// x = x_hi + x_lo with x_hi = x & 0xFFFFE000 (exactly representable in TF32) and x_lo = x - x_hi, and the product is
// x_hi w_hi + x_lo w_hi + x_hi w_lo (three passes, fp32 accumulation in TMEM; measured 7e-6 relative).  The fourth term
// x_lo w_lo is below 2^-22 of the product.
//
// Layout (what makes the implicit GEMM free of any in-kernel gather): activations between tensor-core layers live
// channel-minor, (B, H, W, Cp) with Cp = channels rounded up to 32, as TWO tensors (hi, lo) written by the producing
// layer's epilogue.  One TMA box {32 channels, 18 x, 9 y} lands an input patch as 162 rows of 128 bytes -- exactly the
// K-major SWIZZLE_128B operand layout of the MMA -- and because the MMA unit applies the swizzle to ABSOLUTE
// shared-memory address bits (verified: a descriptor whose start address is offset by any multiple of 128 bytes reads
// the rows shifted by that many pixels, tools/ubench/tc_gemm.cu), the operand of tap (ky, kx) is the SAME buffer with the
// descriptor start advanced by (18 ky + kx) rows.  An output tile is 7 rows x 16 columns enumerated with the input
// pitch (m = 18 y + x, 126 of the 128 MMA rows; x = 16, 17 are scratch columns): 112 useful pixels per 128.
//
// CTA = 4 warps: warp 0 lane 0 issues TMA loads (input patch hi/lo per chunk, double buffered; weights hi/lo per
// (chunk, tap) through a 4-deep ring), warp 1 lane 0 issues the MMAs (12 per stage: 3 passes x 4 k-steps of 8) and frees
// stages with tcgen05.commit, then all four warps read the accumulator (tcgen05.ld 32x32b), add the bias, apply the
// activation, split, and either stage the tile in shared memory for TMA stores into the next layer's (hi, lo) tensors or
// write planar (B, C, H, W) fp32 for the consumers outside the tensor-core chain.
#include "tma.cuh"

#include <algorithm>

namespace b2f {
namespace {
namespace tc {

constexpr int TH = 7, TW = 16, PW = TW + 2;        // output tile; PW = input pitch
constexpr int A_ROWS = (TH + 2) * PW;               // 162 rows of 128 bytes
constexpr int A_BYTES = A_ROWS * 128;               // one of (hi, lo)
constexpr int A_SLOT = (A_BYTES + 4 * 128 + 1023) / 1024 * 1024;   // + the 4 rows the last taps read past the box
#ifndef B2F_TC_NB
#define B2F_TC_NB 2
#endif
#ifndef B2F_TC_AS
#define B2F_TC_AS 1
#endif
constexpr int NB = B2F_TC_NB;                       // weight stages
constexpr int AS = B2F_TC_AS;                       // input-patch stages
// (AS, NB) = (1, 2): 107 KB of shared memory at N = 128 -> TWO CTAs per SM, so that one tile's prologue / epilogue
// (TMEM allocation, first loads, accumulator read-out, stores: 29 % of a tile's residency with one CTA per SM,
// tools/tc_trace.py) runs under the other tile's MMAs; (2, 4) is the one-CTA-per-SM deep-pipeline form.
constexpr int THREADS = 128;

template <int N, bool S2 = false>
struct Cfg {
  static constexpr int B_BYTES = N * 128;           // one of (hi, lo), one (chunk, tap)
  static constexpr int B_SLOT = 2 * B_BYTES;
  static constexpr int A_STAGE = 2 * A_SLOT;
  static constexpr int SMEM_MAIN = AS * A_STAGE + NB * B_SLOT;
  static constexpr int NG = (N + 31) / 32;          // 32-channel output groups
  static constexpr int STAGING = 2 * 2 * (TH * TW * 128);   // two (hi, lo) group buffers, reused round-robin
  static constexpr int SMEM_BYTES = (SMEM_MAIN > STAGING ? SMEM_MAIN : STAGING) + 256 + 1024;
  static constexpr int NACC = S2 ? 4 * N : N;       // stride-2 input gradient: one accumulator per output parity class
  static constexpr int TMEM_COLS = NACC <= 32 ? 32 : (NACC <= 64 ? 64 : (NACC <= 128 ? 128 : (NACC <= 256 ? 256 : 512)));
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(N % 16 == 0 && N >= 16 && N <= 256, "tcgen05.mma M = 128 needs N % 16 == 0");
};

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  // K-major, SWIZZLE_128B (layout type 2 at bits 61-63), SBO = 1024 B (8 rows), LBO = 1 (unused), version 1
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Timeline instrumentation (b2f_debug_tc_trace): 16 clock64 stamps per CTA, only when a buffer is registered.
thread_local unsigned long long* g_tc_trace = nullptr;
#define TC_STAMP(k)                                                                                              \
  do {                                                                                                           \
    if (a.trace) a.trace[(size_t)(blockIdx.x + gridDim.x * blockIdx.y) * 16 + (k)] = clock64();                  \
  } while (0)

struct Args {
  unsigned long long* trace;
  const float* bias;      // [Cout] or NULL
  float* out_planar;      // optional (B, Cout, H, W) fp32, batch stride pbs
  int64_t pbs;
  int nchunk;             // Cin_p / 32
  int Cout, H, W, tiles_x, tiles_y;
  float slope;
  int store_split;        // TMA-store (hi, lo) channel-minor tensors
  // input-gradient use (b2f_conv3x3_tc_backward_data): the activation derivative comes from `mask`, the planar forward
  // output of the layer whose input gradient this is (factor 1 where mask > 0, `slope` elsewhere), not from the
  // result's own sign
  const float* mask;
  int64_t mbs;
  // the same derivative from the channel-minor HI half of that layer's output, (B, H, W, mcp): hi keeps the sign of every
  // normal float, a thread's 32 channels are one 128-byte line instead of 32 strided words (MASK == 2)
  const float* mask_hi;
  int mcp;
  // output-channel slice of a wider convolution (the first decoder layer's input gradient has 196 .. 356 channels): the
  // weight rows start at n0 (out_planar / bias / mask are passed already offset), and the planar result may be ADDED
  int n0, accumulate;
  // S2 (input gradient of a STRIDE-2 convolution): H, W are the low-resolution (output-gradient) size the tiles walk,
  // H2 x W2 the input-gradient plane the four parity classes are written to
  int H2, W2;
};

// MASK: the input-gradient form (the activation derivative comes from a.mask (1, planar) or a.mask_hi (2, channel-minor));
// a compile-time switch -- as a run-time branch inside the unrolled epilogue it cost the forward 10 % (0.307 -> 0.341 ms
// on the level-3 128 -> 128 layer)
// S2: the input gradient of a stride-2 convolution without the 4x zero-inserted detour.  gin[2y + py, 2x + px] only
// receives the taps with ky = py + 1 (mod 2), kx = px + 1 (mod 2), read at g[y + (ky == 0), x + (kx == 0)]: the nine taps
// of the transposed weights are still nine descriptor offsets into ONE low-resolution patch of the output gradient, but
// they accumulate into FOUR TMEM accumulators (1 + 2 + 2 + 4 taps), one per parity class of the output pixel, and the
// epilogue writes each class to its own pixels of the (H2, W2) plane.
template <int N, int MASK, bool S2 = false>
__global__ void __launch_bounds__(THREADS, (AS == 1 ? 2 : 1))
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                  const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                  const __grid_constant__ CUtensorMap tm_oh, const __grid_constant__ CUtensorMap tm_ol, const Args a) {
  using cfg = Cfg<N, S2>;
  static_assert(!S2 || MASK == 0, "the stride-2 input gradient has no activation mask");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_buf = smem;                                  // [2 stages][hi, lo][A_SLOT]
  uint8_t* b_buf = smem + AS * cfg::A_STAGE;              // [NB][hi, lo][B_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (cfg::SMEM_MAIN > cfg::STAGING ? cfg::SMEM_MAIN : cfg::STAGING));
  uint64_t* a_full = bars;                 // [AS] (two slots reserved)
  uint64_t* a_empty = bars + 2;            // [AS]
  uint64_t* b_full = bars + 4;             // [NB]
  uint64_t* b_empty = bars + 4 + NB;       // [NB]
  uint64_t* acc_full = bars + 4 + 2 * NB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * NB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x, b = blockIdx.y;
  const int x0 = tx * TW, y0 = ty * TH;
  // blockIdx.z = slice of N output columns of a wider convolution (the first decoder layer's input gradient has 196 ..
  // 356 channels; at the coarse levels, where a launch has fewer tiles than the machine has SMs, every layer is cut into
  // 32- or 64-column slices so that the serial MMA chain of a CTA is shorter and there are more CTAs)
  const int nz = blockIdx.z * N;

  if (threadIdx.x == 0) {
    TC_STAMP(0);
    if (a.trace) {
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      a.trace[(size_t)(blockIdx.x + gridDim.x * blockIdx.y) * 16 + 9] = sm;
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0 && lane == 0) {
    // ---- TMA producer, weights: one (chunk, tap) stage = hi + lo rows of all N output channels ----
    const int nstage = a.nchunk * 9;
    for (int bs = 0; bs < nstage; ++bs) {
      const int s = bs % NB, c = bs / 9, t = bs % 9;
      if (bs >= NB) mbar_wait(&b_empty[s], ((bs / NB) - 1) & 1);
      mbar_arrive_expect_tx(&b_full[s], 2 * cfg::B_BYTES);
      tma_load_4d(b_buf + s * cfg::B_SLOT, &tm_wh, c * 32, a.n0 + nz, t, 0, &b_full[s]);
      tma_load_4d(b_buf + s * cfg::B_SLOT + cfg::B_BYTES, &tm_wl, c * 32, a.n0 + nz, t, 0, &b_full[s]);
    }
  } else if (warp == 2 && lane == 0) {
    // ---- TMA producer, input patches: its own thread, so that the patch of chunk c + 1 is requested the moment
    // chunk c - 1 retires (a whole chunk of MMAs ahead) instead of behind the weight ring's back-pressure ----
    for (int c = 0; c < a.nchunk; ++c) {
      const int as = c % AS;
      if (c >= AS) mbar_wait(&a_empty[as], ((c / AS) - 1) & 1);
      mbar_arrive_expect_tx(&a_full[as], 2 * A_BYTES);
      tma_load_4d(a_buf + as * cfg::A_STAGE, &tm_xh, c * 32, x0 - 1, y0 - 1, b, &a_full[as]);
      tma_load_4d(a_buf + as * cfg::A_STAGE + A_SLOT, &tm_xl, c * 32, x0 - 1, y0 - 1, b, &a_full[as]);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer ----
    // instruction descriptor: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17,
    // M >> 4 at bit 24
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int bs = 0;
    uint32_t acc = 0;
    uint32_t seen = 0;       // S2: parity classes whose accumulator has been written
    for (int c = 0; c < a.nchunk; ++c) {
      const int as = c % AS;
      mbar_wait(&a_full[as], (c / AS) & 1);
      if (c == 0) TC_STAMP(2);
      const uint32_t ah = smem_u32(a_buf + as * cfg::A_STAGE), al = ah + A_SLOT;
      for (int t = 0; t < 9; ++t, ++bs) {
        const int s = bs % NB;
        mbar_wait(&b_full[s], (bs / NB) & 1);
        if (bs == 0) TC_STAMP(3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t shift = (uint32_t)((t / 3) * PW + (t % 3)) * 128u;
        uint32_t dcol = 0;
        if (S2) {
          // stage t holds the mirrored tap: ky = 2 - t / 3, kx = 2 - t % 3
          const int ky = 2 - t / 3, kx = 2 - t % 3;
          const int cls = ((ky + 1) & 1) * 2 + ((kx + 1) & 1);
          shift = (uint32_t)((1 + (ky == 0)) * PW + 1 + (kx == 0)) * 128u;
          dcol = (uint32_t)(cls * N);
          acc = (seen >> cls) & 1u;
          seen |= 1u << cls;
        }
        const uint32_t bh = smem_u32(b_buf + s * cfg::B_SLOT), bl = bh + cfg::B_BYTES;
#pragma unroll
        for (int p = 0; p < 3; ++p) {            // hi * hi, lo * hi, hi * lo
          const uint32_t pa = (p == 1 ? al : ah) + shift, pb = (p == 2 ? bl : bh);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_tf32(tmem + dcol, smem_desc(pa + 32u * k), smem_desc(pb + 32u * k), idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&b_empty[s]);               // the weight stage is free once these MMAs have read it
      }
      umma_commit(&a_empty[as]);
    }
    umma_commit(acc_full);
    TC_STAMP(4);
  }
  __syncwarp();

  // ---- epilogue: all four warps; warp w owns TMEM lanes 32 w .. 32 w + 31 = tile rows m ----
  mbar_wait(acc_full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) TC_STAMP(5);
  const int m = 32 * warp + lane;
  const int ry = m / PW, rx = m % PW;
  const bool valid = ry < TH && rx < TW;
  const int y = y0 + ry, x = x0 + rx;
  const bool inside = valid && y < a.H && x < a.W;
  const int prow = ry * TW + rx;                          // row of the staged (dense 7 x 16) tile
  const uint32_t stage0 = smem_u32(smem);
  if (S2) {
    // four parity classes x NG groups, planar only: class (py, px) of tile pixel (y, x) is pixel (2y + py, 2x + px)
#pragma unroll 1
    for (int q = 0; q < 4 * cfg::NG; ++q) {
      const int cls = q / cfg::NG, g = q % cfg::NG;
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(cls * N + 32 * g), v);
      const int Y = 2 * y + (cls >> 1), X = 2 * x + (cls & 1);
      if (valid && Y < a.H2 && X < a.W2) {
        float* o = a.out_planar + (size_t)b * a.pbs + (size_t)Y * a.W2 + X;
        const size_t cs = (size_t)a.H2 * a.W2;
        if (a.accumulate) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nz + 32 * g + j < a.Cout) atomicAdd(o + (size_t)(nz + 32 * g + j) * cs, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nz + 32 * g + j < a.Cout) o[(size_t)(nz + 32 * g + j) * cs] = __uint_as_float(v[j]);
        }
      }
    }
  }
#pragma unroll 1
  for (int g = 0; g < (S2 ? 0 : cfg::NG); ++g) {
    uint32_t v[32];
    float mk2[MASK == 2 ? 32 : 1];
    if (MASK == 2) {      // requested before the accumulator read so that the two latencies overlap
      const float4* mp = reinterpret_cast<const float4*>(a.mask_hi + (((size_t)b * a.H + (inside ? y : 0)) * a.W + (inside ? x : 0)) * a.mcp +
                                                         a.n0 + nz + 32 * g);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t4 = inside ? __ldg(mp + q) : make_float4(1.f, 1.f, 1.f, 1.f);
        mk2[4 * q] = t4.x; mk2[4 * q + 1] = t4.y; mk2[4 * q + 2] = t4.z; mk2[4 * q + 3] = t4.w;
      }
    }
    tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(32 * g), v);
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = nz + 32 * g + j;
      float t = __uint_as_float(v[j]) + ((a.bias && n < a.Cout) ? __ldg(a.bias + n) : 0.f);
      if (MASK == 2) {
        t = mk2[j] > 0.f ? t : t * a.slope;
      } else if (MASK == 1) {
        const float mk = (inside && n < a.Cout) ? __ldg(a.mask + (size_t)b * a.mbs + ((size_t)n * a.H + y) * a.W + x) : 1.f;
        t = mk > 0.f ? t : t * a.slope;
      } else {
        t = t > 0.f ? t : t * a.slope;
      }
      f[j] = n < a.Cout ? t : 0.f;
    }
    if (a.out_planar && inside) {
      float* o = a.out_planar + (size_t)b * a.pbs + (size_t)y * a.W + x;
      if (a.accumulate) {
        // one writer per element and call: a reduction without a return value (RED) instead of load + add + store,
        // so that the thread does not wait for 32 strided loads
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nz + 32 * g + j < a.Cout) atomicAdd(o + (size_t)(nz + 32 * g + j) * a.H * a.W, f[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nz + 32 * g + j < a.Cout) o[(size_t)(nz + 32 * g + j) * a.H * a.W] = f[j];
      }
    }
    if (a.store_split) {
      // two (hi, lo) staging buffers: group g reuses the buffer of group g - 2 once its stores have read it
      if (g >= 2) {
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
      }
      const uint32_t rh0 = stage0 + (uint32_t)((g & 1) * 2 * (TH * TW * 128));
      if (valid) {
        const uint32_t rh = rh0 + (uint32_t)(prow * 128);
        const uint32_t rl = rh + TH * TW * 128;
        const uint32_t sw = (uint32_t)(prow & 7);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 hi, lo;
          hi.x = __uint_as_float(__float_as_uint(f[4 * q]) & 0xFFFFE000u);
          hi.y = __uint_as_float(__float_as_uint(f[4 * q + 1]) & 0xFFFFE000u);
          hi.z = __uint_as_float(__float_as_uint(f[4 * q + 2]) & 0xFFFFE000u);
          hi.w = __uint_as_float(__float_as_uint(f[4 * q + 3]) & 0xFFFFE000u);
          lo.x = f[4 * q] - hi.x; lo.y = f[4 * q + 1] - hi.y; lo.z = f[4 * q + 2] - hi.z; lo.w = f[4 * q + 3] - hi.w;
          sts128(rh + 16u * ((uint32_t)q ^ sw), hi);
          sts128(rl + 16u * ((uint32_t)q ^ sw), lo);
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        tma_store_4d_addr(rh0, &tm_oh, nz + 32 * g, x0, y0, b);
        tma_store_4d_addr(rh0 + TH * TW * 128, &tm_ol, nz + 32 * g, x0, y0, b);
        tma_store_commit();
      }
    }
  }
  if (threadIdx.x == 0) TC_STAMP(6);
  if (a.store_split && threadIdx.x == 0) {
    tma_store_wait_read();
    TC_STAMP(7);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cfg::TMEM_COLS) : "memory");
  if (threadIdx.x == 0) TC_STAMP(8);
}

// planar (B, C, H, W) [batch stride xbs] -> channel-minor (B, H, W, Cp) hi / lo.  A block transposes 32 channels x 128
// pixels through shared memory: 16 coalesced loads per thread in flight before the barrier, then 16 pixels x (hi, lo)
// of 128-byte channel rows per warp (with 32 pixels per block -- 4 loads per thread -- the kernel ran at 3.7 TB/s).
constexpr int SPLIT_PX = 128;
__global__ void __launch_bounds__(256) split_from_planar_kernel(const float* __restrict__ x, int64_t xbs, float* __restrict__ hi,
                                                               float* __restrict__ lo, int C, int Cp, int64_t hw) {
  __shared__ float tile[32][SPLIT_PX + 1];
  const int64_t p0 = (int64_t)blockIdx.x * SPLIT_PX;
  const int c0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8 threads
  float v[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const float* src = x + (size_t)b * xbs + (size_t)c * hw + p0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t p = p0 + tx + 32 * q;
      v[k][q] = (c < C && p < hw) ? __ldg(src + tx + 32 * q) : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) tile[ty + 8 * k][tx + 32 * q] = v[k][q];
  __syncthreads();
#pragma unroll 4
  for (int i = ty; i < SPLIT_PX; i += 8) {
    const int64_t p = p0 + i;
    if (p >= hw) break;
    const float t = tile[tx][i];
    const float h = __uint_as_float(__float_as_uint(t) & 0xFFFFE000u);
    const size_t o = ((size_t)b * hw + p) * Cp + c0 + tx;
    hi[o] = h;
    lo[o] = t - h;
  }
}

// Torch (Cout, Cin, 3, 3) -> [9 taps][Cout][Cin_p] hi / lo (K-major rows of the B operand)
__global__ void pack_tc_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int Cout,
                                       int Cin, int CinP) {
  const int total = 9 * Cout * CinP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % CinP, n = (i / CinP) % Cout, t = i / (CinP * Cout);
    const float v = ci < Cin ? w[((size_t)n * Cin + ci) * 9 + t] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

// packed FFMA layout [Cin * 9][CoutP] (conv.cu) -> tensor-core layout [9][N][Kp] hi / lo, on the device: the training
// path re-packs the decoders' weights at the start of every step (Adam updates the flat packed parameters).
// transpose = 0: the forward operand, N = Cout, K = input channels (zero beyond Cin up to Kp);
// transpose = 1: the input-gradient operand, N = Cin, K = output channels, taps mirrored (w'[ci][co][t] = w[co][ci][8 - t]).
__global__ void pack_tc_from_packed_kernel(const float* __restrict__ wp, float* __restrict__ hi, float* __restrict__ lo, int Cout,
                                           int Cin, int CoutP, int N, int Kp, int transpose) {
  const int total = 9 * N * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % Kp, n = (i / Kp) % N, t = i / (Kp * N);
    float v = 0.f;
    if (!transpose) {
      if (k < Cin && n < Cout) v = wp[((size_t)k * 9 + t) * CoutP + n];
    } else {
      if (k < Cout && n < Cin) v = wp[((size_t)n * 9 + (8 - t)) * CoutP + k];
    }
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

// blockIdx.y = job (b2f_pack_job in device memory), blockIdx.x strides over the job's elements
struct PackJob {
  const float* wp;
  float* hi;
  float* lo;
  int32_t Cout, Cin, K, transpose;
};
__global__ void pack_tc_from_packed_batch_kernel(const PackJob* __restrict__ jobs) {
  const PackJob j = jobs[blockIdx.y];
  const int Kp = (j.K + 31) / 32 * 32, N = j.transpose ? j.Cin : j.Cout, CoutP = (j.Cout + 63) / 64 * 64;
  const int total = 9 * N * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % Kp, n = (i / Kp) % N, t = i / (Kp * N);
    float v = 0.f;
    if (!j.transpose) {
      if (k < j.Cin && n < j.Cout) v = j.wp[((size_t)k * 9 + t) * CoutP + n];
    } else {
      if (k < j.Cout && n < j.Cin) v = j.wp[((size_t)n * 9 + (8 - t)) * CoutP + k];
    }
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    j.hi[i] = h;
    j.lo[i] = v - h;
  }
}

template <int N, int MASK, bool S2 = false>
int launch_tc(const float* xh, const float* xl, const float* wh, const float* wl, float* oh, float* ol, const Args& a0, int B,
              int CinP, int CoutP, cudaStream_t st, int wrows_total = 0, int nslices = 1) {
  using cfg = Cfg<N, S2>;
  Args a = a0;
  CUtensorMap txh, txl, twh, twl, toh, tol;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)CinP, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)CinP, (uint64_t)CinP * a.W, (uint64_t)CinP * a.W * a.H};
    const uint32_t box[4] = {32, (uint32_t)PW, (uint32_t)(TH + 2), 1};
    if ((rc = make_tmap4(&txh, xh, dims, str, box, true))) return rc;
    if ((rc = make_tmap4(&txl, xl, dims, str, box, true))) return rc;
  }
  {
    const uint64_t wrows = (uint64_t)(wrows_total > 0 ? wrows_total : a.Cout);
    const uint64_t dims[4] = {(uint64_t)CinP, wrows, 9, 1};
    const uint64_t str[3] = {(uint64_t)CinP, (uint64_t)CinP * wrows, (uint64_t)CinP * wrows * 9};
    const uint32_t box[4] = {32, (uint32_t)N, 1, 1};
    if ((rc = make_tmap4(&twh, wh, dims, str, box, true))) return rc;
    if ((rc = make_tmap4(&twl, wl, dims, str, box, true))) return rc;
  }
  if (a.store_split) {
    const uint64_t dims[4] = {(uint64_t)CoutP, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)CoutP, (uint64_t)CoutP * a.W, (uint64_t)CoutP * a.W * a.H};
    const uint32_t box[4] = {32, (uint32_t)TW, (uint32_t)TH, 1};
    if ((rc = make_tmap4(&toh, oh, dims, str, box, true))) return rc;
    if ((rc = make_tmap4(&tol, ol, dims, str, box, true))) return rc;
  } else {
    toh = txh;
    tol = txl;
  }
  auto kern = conv3x3_tc_kernel<N, MASK, S2>;
  static thread_local int attr_dev = -1;
  int dev = 0;
  B2F_CUDA_TRY(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    B2F_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    attr_dev = dev;
  }
  a.tiles_x = (a.W + TW - 1) / TW;
  a.tiles_y = (a.H + TH - 1) / TH;
  dim3 grid(a.tiles_x * a.tiles_y, B, nslices);
  kern<<<grid, THREADS, cfg::SMEM_BYTES, st>>>(txh, txl, twh, twl, toh, tol, a);
  B2F_CHECK_LAUNCH("conv3x3_tc_kernel");
  return B2F_OK;
}

}  // namespace tc
}  // namespace
// the registered timeline buffer, shared with the tensor-core cost volume (costvol_tc.cu)
unsigned long long* tc_trace_buffer() { return tc::g_tc_trace; }
}  // namespace b2f

using namespace b2f;

extern "C" int b2f_debug_tc_trace(unsigned long long* device_buffer) {
  tc::g_tc_trace = device_buffer;
  return B2F_OK;
}

extern "C" int64_t b2f_conv3x3_tc_packed_floats(int Cin, int Cout) {
  if (Cin <= 0 || Cout <= 0) return 0;
  return (int64_t)9 * Cout * ((Cin + 31) / 32 * 32);
}

extern "C" int b2f_conv3x3_tc_pack_weights(const float* w_torch, float* w_hi, float* w_lo, int Cout, int Cin,
                                           b2f_stream_t stream) {
  if (!w_torch || !w_hi || !w_lo || Cout <= 0 || Cin <= 0) return fail(B2F_EINVAL, "conv3x3_tc_pack_weights: bad argument");
  const int CinP = (Cin + 31) / 32 * 32;
  const int total = 9 * Cout * CinP;
  tc::pack_tc_weights_kernel<<<std::max(1, std::min((total + 255) / 256, num_sms() * 8)), 256, 0,
                               reinterpret_cast<cudaStream_t>(stream)>>>(w_torch, w_hi, w_lo, Cout, Cin, CinP);
  B2F_CHECK_LAUNCH("pack_tc_weights_kernel");
  return B2F_OK;
}

extern "C" int b2f_conv3x3_tc_pack_from_packed(const float* w_packed, float* w_hi, float* w_lo, int Cout, int Cin, int K,
                                               int transpose, b2f_stream_t stream) {
  if (!w_packed || !w_hi || !w_lo || Cout <= 0 || Cin <= 0) return fail(B2F_EINVAL, "conv3x3_tc_pack_from_packed: bad argument");
  const int kmin = transpose ? Cout : Cin;
  if (K < kmin) return fail(B2F_EINVAL, "conv3x3_tc_pack_from_packed: K = %d smaller than the %d channels of the weights", K, kmin);
  const int Kp = (K + 31) / 32 * 32, N = transpose ? Cin : Cout, CoutP = (Cout + 63) / 64 * 64;
  const int total = 9 * N * Kp;
  tc::pack_tc_from_packed_kernel<<<std::max(1, std::min((total + 255) / 256, num_sms() * 8)), 256, 0,
                                   reinterpret_cast<cudaStream_t>(stream)>>>(w_packed, w_hi, w_lo, Cout, Cin, CoutP, N, Kp, transpose);
  B2F_CHECK_LAUNCH("pack_tc_from_packed_kernel");
  return B2F_OK;
}

static_assert(sizeof(b2f_pack_job) == sizeof(tc::PackJob), "b2f_pack_job layout");
extern "C" int b2f_conv3x3_tc_pack_from_packed_batch(const b2f_pack_job* jobs_device, int njobs, b2f_stream_t stream) {
  if (njobs < 0 || (njobs > 0 && !jobs_device)) return fail(B2F_EINVAL, "conv3x3_tc_pack_from_packed_batch: bad argument");
  if (njobs > 65535) return fail(B2F_EINVAL, "conv3x3_tc_pack_from_packed_batch: more than 65535 jobs");
  if (njobs == 0) return B2F_OK;
  tc::pack_tc_from_packed_batch_kernel<<<dim3(16, njobs), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const tc::PackJob*>(jobs_device));
  B2F_CHECK_LAUNCH("pack_tc_from_packed_batch_kernel");
  return B2F_OK;
}

extern "C" int b2f_nhwc_split_from_bdhw(const float* x, int64_t x_batch_stride, float* hi, float* lo, int B, int C, int H, int W,
                                        b2f_stream_t stream) {
  if (!x || !hi || !lo || B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "nhwc_split_from_bdhw: bad argument");
  if (B == 0) return B2F_OK;
  const int Cp = (C + 31) / 32 * 32;
  const int64_t hw = (int64_t)H * W;
  const int64_t xbs = x_batch_stride ? x_batch_stride : (int64_t)C * hw;
  if (B > 65535) return fail(B2F_EINVAL, "nhwc_split_from_bdhw: B > 65535");
  dim3 grid((unsigned)((hw + tc::SPLIT_PX - 1) / tc::SPLIT_PX), Cp / 32, B);
  tc::split_from_planar_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, xbs, hi, lo, C, Cp, hw);
  B2F_CHECK_LAUNCH("split_from_planar_kernel");
  return B2F_OK;
}

// Slice width (the kernel's N) and number of slices for `width` output columns over `ctas` = tiles x images CTAs per
// slice.  A machine-filling launch takes the widest slices (fewest re-reads of the input patch): <= 128 columns in one,
// more in ceil(width / 128) balanced ones.  A launch with fewer CTAs than SMs is latency-bound on the serial MMA chain
// of its CTAs (a 128 -> 128 layer takes 36 us at 5 x 10, 10 x 20 and 20 x 40 alike): 32-column slices.
static void tc_slices(int width, int ctas, int* ns, int* nslices) {
  const int w32 = (width + 31) / 32 * 32;
  int n;
  if (w32 <= 32) n = 32;
  else if (ctas < 100) n = 32;      // (64-column slices at 240 CTAs measured 25 % SLOWER than one 128-column slice)
  else {
    const int z = (w32 + 127) / 128;
    n = ((w32 + z - 1) / z + 31) / 32 * 32;
  }
  if (n > w32) n = w32;
  *ns = n;
  *nslices = (w32 + n - 1) / n;
}

template <int MASK>
static int tc_dispatch(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo, float* out_hi, float* out_lo,
                       const tc::Args& a, int B, int Cin, int N, int nslices, int width, cudaStream_t st, const char* who,
                       int wrows_total) {
  const int CinP = (Cin + 31) / 32 * 32, CoutP = (width + 31) / 32 * 32;
  switch (N) {
    case 32: return tc::launch_tc<32, MASK>(x_hi, x_lo, w_hi, w_lo, out_hi, out_lo, a, B, CinP, CoutP, st, wrows_total, nslices);
    case 64: return tc::launch_tc<64, MASK>(x_hi, x_lo, w_hi, w_lo, out_hi, out_lo, a, B, CinP, CoutP, st, wrows_total, nslices);
    case 96: return tc::launch_tc<96, MASK>(x_hi, x_lo, w_hi, w_lo, out_hi, out_lo, a, B, CinP, CoutP, st, wrows_total, nslices);
    case 128: return tc::launch_tc<128, MASK>(x_hi, x_lo, w_hi, w_lo, out_hi, out_lo, a, B, CinP, CoutP, st, wrows_total, nslices);
    default: return fail(B2F_EUNSUPPORTED, "%s: slice width %d", who, N);
  }
}

// Input gradient of a stride-1 3x3 convolution on the tensor cores: the forward kernel on the transposed, mirrored
// weights (b2f_conv3x3_tc_pack_from_packed, transpose = 1) over the channel-minor (hi, lo) OUTPUT gradient, with the
// LeakyReLU derivative of the layer below taken from its planar forward output `act`.
extern "C" int b2f_conv3x3_tc_backward_data(const float* g_hi, const float* g_lo, const float* wt_hi, const float* wt_lo,
                                            const float* act, int64_t act_batch_stride, const float* act_hi, float* gin_hi, float* gin_lo,
                                            float* gin_planar, int64_t gin_planar_batch_stride, int B, int Cout, int H, int W,
                                            int Cin, float leaky_slope, int accumulate, b2f_stream_t stream) {
  if (!g_hi || !g_lo || !wt_hi || !wt_lo) return fail(B2F_EINVAL, "conv3x3_tc_backward_data: NULL gradient / weights");
  if ((gin_hi == nullptr) != (gin_lo == nullptr)) return fail(B2F_EINVAL, "conv3x3_tc_backward_data: gin_hi and gin_lo go together");
  if (!gin_hi && !gin_planar) return fail(B2F_EINVAL, "conv3x3_tc_backward_data: no output");
  if (B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || B > 65535) return fail(B2F_EINVAL, "conv3x3_tc_backward_data: bad size");
  if (!aligned16(g_hi) || !aligned16(g_lo) || !aligned16(wt_hi) || !aligned16(wt_lo) || (gin_hi && (!aligned16(gin_hi) || !aligned16(gin_lo))))
    return fail(B2F_EALIGN, "conv3x3_tc_backward_data: operands must be 16-byte aligned");
  if (get_encode_fn() == nullptr) return fail(B2F_EUNSUPPORTED, "conv3x3_tc_backward_data: cuTensorMapEncodeTiled not available");
  if (accumulate && !gin_planar) return fail(B2F_EINVAL, "conv3x3_tc_backward_data: accumulate needs the planar output");
  if (act && act_hi) return fail(B2F_EINVAL, "conv3x3_tc_backward_data: act (planar) and act_hi (channel-minor) are alternatives");
  if (act_hi && !aligned16(act_hi)) return fail(B2F_EALIGN, "conv3x3_tc_backward_data: act_hi must be 16-byte aligned");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t pbs = gin_planar_batch_stride ? gin_planar_batch_stride : (int64_t)Cin * H * W;
  const int64_t mbs = act_batch_stride ? act_batch_stride : (int64_t)Cin * H * W;
  int N, nslices;
  tc_slices(Cin, ((W + tc::TW - 1) / tc::TW) * ((H + tc::TH - 1) / tc::TH) * B, &N, &nslices);
  tc::Args a{};
  a.trace = tc::g_tc_trace;
  a.bias = nullptr;
  a.out_planar = gin_planar;
  a.pbs = pbs;
  a.nchunk = ((Cout + 31) / 32 * 32) / 32;      // K = the forward layer's output channels
  a.Cout = Cin; a.H = H; a.W = W;               // valid output columns of the launch
  a.slope = (act || act_hi) ? leaky_slope : 1.f;
  a.store_split = gin_hi != nullptr;
  a.mask = act;
  a.mbs = mbs;
  a.mask_hi = act_hi;
  a.mcp = (Cin + 31) / 32 * 32;
  a.n0 = 0;
  a.accumulate = accumulate;
  return act_hi ? tc_dispatch<2>(g_hi, g_lo, wt_hi, wt_lo, gin_hi, gin_lo, a, B, Cout, N, nslices, Cin, st, "conv3x3_tc_backward_data", Cin)
         : act  ? tc_dispatch<1>(g_hi, g_lo, wt_hi, wt_lo, gin_hi, gin_lo, a, B, Cout, N, nslices, Cin, st, "conv3x3_tc_backward_data", Cin)
                : tc_dispatch<0>(g_hi, g_lo, wt_hi, wt_lo, gin_hi, gin_lo, a, B, Cout, N, nslices, Cin, st, "conv3x3_tc_backward_data", Cin);
}

// Input gradient of a STRIDE-2 3x3 convolution (the down-sampling layers of the feature pyramid) on the tensor cores:
// four parity-class accumulators over one low-resolution patch of the output gradient (kernel comment, S2).
extern "C" int b2f_conv3x3_tc_backward_data_s2(const float* g_hi, const float* g_lo, const float* wt_hi, const float* wt_lo,
                                               float* gin_planar, int64_t gin_planar_batch_stride, int B, int Cout, int Ho,
                                               int Wo, int Cin, int H, int W, int accumulate, b2f_stream_t stream) {
  if (!g_hi || !g_lo || !wt_hi || !wt_lo || !gin_planar) return fail(B2F_EINVAL, "conv3x3_tc_backward_data_s2: NULL operand");
  if (B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || B > 65535) return fail(B2F_EINVAL, "conv3x3_tc_backward_data_s2: bad size");
  if (Ho != (H - 1) / 2 + 1 || Wo != (W - 1) / 2 + 1)
    return fail(B2F_EINVAL, "conv3x3_tc_backward_data_s2: %d x %d is not the stride-2 output size of %d x %d", Ho, Wo, H, W);
  if (Cin > 128) return fail(B2F_EUNSUPPORTED, "conv3x3_tc_backward_data_s2: Cin = %d > 128", Cin);
  if (!aligned16(g_hi) || !aligned16(g_lo) || !aligned16(wt_hi) || !aligned16(wt_lo) || !aligned4(gin_planar))
    return fail(B2F_EALIGN, "conv3x3_tc_backward_data_s2: misaligned operand");
  if (get_encode_fn() == nullptr) return fail(B2F_EUNSUPPORTED, "conv3x3_tc_backward_data_s2: cuTensorMapEncodeTiled not available");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int KP = (Cout + 31) / 32 * 32;
  int N, nslices;
  tc_slices(Cin, ((Wo + tc::TW - 1) / tc::TW) * ((Ho + tc::TH - 1) / tc::TH) * B, &N, &nslices);
  tc::Args a{};
  a.trace = tc::g_tc_trace;
  a.out_planar = gin_planar;
  a.pbs = gin_planar_batch_stride ? gin_planar_batch_stride : (int64_t)Cin * H * W;
  a.nchunk = KP / 32;
  a.Cout = Cin; a.H = Ho; a.W = Wo; a.H2 = H; a.W2 = W;
  a.slope = 1.f;
  a.accumulate = accumulate;
  const int NP = (Cin + 31) / 32 * 32;
  switch (N) {
    case 32: return tc::launch_tc<32, 0, true>(g_hi, g_lo, wt_hi, wt_lo, nullptr, nullptr, a, B, KP, NP, st, Cin, nslices);
    case 64: return tc::launch_tc<64, 0, true>(g_hi, g_lo, wt_hi, wt_lo, nullptr, nullptr, a, B, KP, NP, st, Cin, nslices);
    case 96: return tc::launch_tc<96, 0, true>(g_hi, g_lo, wt_hi, wt_lo, nullptr, nullptr, a, B, KP, NP, st, Cin, nslices);
    default: return tc::launch_tc<128, 0, true>(g_hi, g_lo, wt_hi, wt_lo, nullptr, nullptr, a, B, KP, NP, st, Cin, nslices);
  }
}

extern "C" int b2f_conv3x3_tc_forward(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo,
                                      const float* bias, float* out_hi, float* out_lo, float* out_planar,
                                      int64_t out_planar_batch_stride, int B, int Cin, int H, int W, int Cout,
                                      float leaky_slope, b2f_stream_t stream) {
  if (!x_hi || !x_lo || !w_hi || !w_lo) return fail(B2F_EINVAL, "conv3x3_tc_forward: NULL input / weights");
  if ((out_hi == nullptr) != (out_lo == nullptr)) return fail(B2F_EINVAL, "conv3x3_tc_forward: out_hi and out_lo go together");
  if (!out_hi && !out_planar) return fail(B2F_EINVAL, "conv3x3_tc_forward: no output");
  if (B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "conv3x3_tc_forward: bad size");
  if (B > 65535) return fail(B2F_EINVAL, "conv3x3_tc_forward: B > 65535");
  if (!aligned16(x_hi) || !aligned16(x_lo) || !aligned16(w_hi) || !aligned16(w_lo) || (out_hi && (!aligned16(out_hi) || !aligned16(out_lo))))
    return fail(B2F_EALIGN, "conv3x3_tc_forward: operands must be 16-byte aligned");
  if (get_encode_fn() == nullptr) return fail(B2F_EUNSUPPORTED, "conv3x3_tc_forward: cuTensorMapEncodeTiled not available");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int CinP = (Cin + 31) / 32 * 32;
  (void)CinP;
  const int64_t pbs = out_planar_batch_stride ? out_planar_batch_stride : (int64_t)Cout * H * W;
  // output-channel slices in ONE launch (blockIdx.z): tc_slices
  int N, nslices;
  tc_slices(Cout, ((W + tc::TW - 1) / tc::TW) * ((H + tc::TH - 1) / tc::TH) * B, &N, &nslices);
  tc::Args a{};
  a.trace = tc::g_tc_trace;
  a.bias = bias;
  a.out_planar = out_planar;
  a.pbs = pbs;
  a.nchunk = ((Cin + 31) / 32 * 32) / 32;
  a.Cout = Cout; a.H = H; a.W = W;
  a.slope = leaky_slope;
  a.store_split = out_hi != nullptr;
  a.n0 = 0;
  return tc_dispatch<0>(x_hi, x_lo, w_hi, w_lo, out_hi, out_lo, a, B, Cin, N, nslices, Cout, st, "conv3x3_tc_forward", Cout);
}
