// nn.CostVolMulti forward (F = 2, win = 9) on the 5th-generation tensor cores (tcgen05 / TMEM) for sm_100a.
//
// models/CostVolMulti.lua:49-110 computes, per pixel p and displacement q in the 9 x 9 window,
//     out[q, p] = 1/C sum_c ref[c, p] * frame[c, p - s q].
// Per 8 x 16 pixel tile that is a BAND of the dense product
//     D[m, n] = sum_c ref[c, m] * frame[c, n],     m = 16 r + x  (128 tile pixels),   n = 24 hr + hc  (16 x 24 halo)
// of which a pixel uses the 9 x 9 block (hr, hc) = (r + dy, x + dx), dy, dx = 0..8: 81 of 384 columns (21 %).  The
// tensor core does not mind: M = 128, N = 384, K = C as three kind::tf32 passes of the (hi, lo) split (x_hi = x &
// 0xFFFFE000 is exact in TF32, x_lo = x - x_hi; hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM; 2e-6 .. 5e-6
// relative against the float64 checker) is 36 MMAs = 2 304 nominal / ~3 200 measured cycles per tile against
// ~8 300 cycles of the FFMA2 kernel (costvol.cu, cvf::costvol_fwd_tma).  Everything AROUND the MMA is what decides,
// so each piece below is there because a clock64 timeline (tools/cvt_trace.py) showed the previous form stalling:
//   * feed: no TMA -- a tile needs C x (8 + 16) box rows and the TMA unit delivers one box row per ~8 cycles per SM
//     whatever its length (tools/ubench/tma_feed.cu): 6 000 cycles per tile.  Two LOADER warps copy the planar maps
//     with 16-byte cp.async (four consecutive x of one channel; out-of-image halo positions and channels >= C are
//     zero-filled by the copy's src-size = "out-of-range terms are dropped", CostVolMulti.lua:77-88) into a raw
//     [group][channel][row] ring and signal with cp.async.mbarrier.arrive.  (Loading into registers with LDG was tried
//     first: the proxy fence the operand stores need is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, which waits for the
//     thread's prefetched loads too -- one L2 round trip per group, 9 400 cycles per tile.  4-byte cp.async: issue
//     bound, ~7 cycles per LDGSTS.)
//   * operands: four CONVERTER warps, one per 128-row operand group (A = the reference tile, B0..B2 = the halo in
//     three N = 128 chunks), read a row's 32 channels from the ring, split, and write the K-major SWIZZLE_128B rows
//     (128 bytes = 32 channels): the transpose planar -> channel-minor and the split are one pass, with one proxy
//     fence per group and warp.  A is double buffered.
//   * accumulators: the tile's three chunks rotate through the FOUR 128-column slots of the 512-column TMEM with a
//     full / empty barrier per slot, so chunk 0 of tile t + 1 is computed while tile t is still being read;
//   * epilogue: eight warps, two per TMEM lane quadrant q (= tile rows 2q, 2q + 1) taking the even / odd ones of its
//     ten halo rows.  tcgen05.ld columns are warp-uniform but a pixel's window starts at its own x, so per halo row a
//     warp loads the 24 columns (next row's loads in flight) and every lane barrel-shifts by x (47 SEL) to its nine
//     values, which leave as 64-byte runs (16 pixels of one output plane).
// Result (B200, level 3, B = 8; tools/time_cv_tc.py, profiles/r02_costvol_tc_vs_ffma2.txt): 53-55 us against 51-52 us
// for the FFMA2 kernel (level 4: 28.7 vs 31.5, level 5: 16.4 vs 18.4, levels 6-7 slower).  43 us without the global
// stores (648 sixty-four-byte segments per tile through the LSU), and ncu shows why it stops there: the shared-memory
// data path carries 2 304 wavefronts of MMA operand fetch + ~2 000 of operand / ring writes and reads per tile -- the
// tensor core is busy a third of the time and the kernel is bound by the same 128 B/clk port as the FFMA2 kernel.
// The local-window correlation is NOT a dense contraction on this machine; the automatic dispatch keeps the FFMA2
// kernel and this one stays selectable (b2f_debug_costvol_path 16) with its parity tests.
// Shapes: any C (zero-padded to a multiple of 32), any H; W % 4 == 0 and 16-byte aligned maps (16-byte copies).
#include "tma.cuh"

#include <algorithm>

namespace b2f {
unsigned long long* tc_trace_buffer();   // conv_tc.cu (b2f_debug_tc_trace)
namespace {
namespace cvt {

constexpr int TH = 8, TW = 16;             // pixel tile: M = 128
constexpr int HC = TW + 8;                 // halo columns (24); halo rows TH + 8 = 16: N = 384
constexpr int A_BYTES = 128 * 128;         // one of (hi, lo): 128 rows x 32 channels x 4 B
constexpr int B_BYTES = 384 * 128;
// A (the reference tile) is double buffered: the converters write tile t + 1's while tile t's last MMAs still read theirs
constexpr int OFF_AH = 0, OFF_AL = A_BYTES, A_STAGE = 2 * A_BYTES, OFF_BH = 2 * A_STAGE, OFF_BL = 2 * A_STAGE + B_BYTES;
constexpr int SMEM_OPS = 2 * A_STAGE + 2 * B_BYTES;   // 160 KB
constexpr int RAW_BYTES = 4 * 32 * 128 * 4;          // four raw groups in flight (64 KB)
constexpr int SMEM_BYTES = SMEM_OPS + RAW_BYTES + 256 + 1024;
constexpr int MMA_WARP = 14;
// warps 0-7 epilogue (two per TMEM lane quadrant), 8-11 converters, 12-13 loaders (two rows per thread), 14 MMA issuer:
// 480 threads -- the register file is handed out per four warps, 16 warps leave 128 registers per thread
constexpr int THREADS = 480;

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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

struct Args {
  const float* ref;
  const float* frm;
  float* out;
  int64_t obs;       // batch stride of out
  int C, H, W, ntx, nty, ntiles;
  float kinv;
  int dbg;           // measurement aids (WRONG results): 1 = no global stores, 2 = no TMEM loads
  unsigned long long* trace;   // b2f_debug_tc_trace: 16 clock64 stamps per (CTA < 8, tile < 8), tools/cvt_trace.py
};
#define CVT_STAMP(tl, k)                                                                              \
  do {                                                                                                \
    if (a.trace && blockIdx.x < 8 && (tl) < 8) a.trace[((size_t)blockIdx.x * 8 + (tl)) * 16 + (k)] = clock64(); \
  } while (0)

// split into (hi, lo) and write the row into both K-major SWIZZLE_128B operand buffers
__device__ __forceinline__ void store_row(const float (&v)[32], uint32_t row_hi, uint32_t row_lo, uint32_t sw) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v[4 * q]) & 0xFFFFE000u);
    hi.y = __uint_as_float(__float_as_uint(v[4 * q + 1]) & 0xFFFFE000u);
    hi.z = __uint_as_float(__float_as_uint(v[4 * q + 2]) & 0xFFFFE000u);
    hi.w = __uint_as_float(__float_as_uint(v[4 * q + 3]) & 0xFFFFE000u);
    lo.x = v[4 * q] - hi.x; lo.y = v[4 * q + 1] - hi.y; lo.z = v[4 * q + 2] - hi.z; lo.w = v[4 * q + 3] - hi.w;
    sts128(row_hi + 16u * ((uint32_t)q ^ sw), hi);
    sts128(row_lo + 16u * ((uint32_t)q ^ sw), lo);
  }
}

template <int SGN>
__global__ void __launch_bounds__(THREADS, 1) costvol_fwd_tc(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* raw = reinterpret_cast<float*>(smem + SMEM_OPS);       // [4 slots][32 channels][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_OPS + RAW_BYTES);
  uint64_t* ops_full = bars;          // [4]: A, B chunk 0..2 written by the converters
  uint64_t* ops_empty = bars + 4;     // [4]: the MMAs that read them have retired
  uint64_t* tmem_empty = bars + 8;    // [4]: the epilogue has read the slot
  uint64_t* acc_full = bars + 12;     // (unused)
  uint64_t* raw_full = bars + 13;     // [4]: the loaders' cp.async of a group have landed
  uint64_t* raw_empty = bars + 17;    // [4]: the converters hold the group in registers
  uint64_t* a_full = bars + 21;       // [2]: A buffer (step & 1) written (ops_full[0] / ops_empty[0] are unused)
  uint64_t* a_empty = bars + 23;      // [2]
  uint64_t* tmem_full = bars + 25;    // [4]: the MMAs into the slot have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkc = (a.C + 31) / 32;
  const int64_t hw = (int64_t)a.H * a.W;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&ops_full[i], 1);
      mbar_init(&ops_empty[i], 1);
      mbar_init(&tmem_empty[i], 8);
      mbar_init(&raw_full[i], 64);
      mbar_init(&raw_empty[i], 1);
      mbar_init(&tmem_full[i], 1);
    }
    mbar_init(acc_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // ---------------- epilogue: warps q and q + 4 share TMEM lane quadrant q (tile rows 2q, 2q + 1) and take the even /
    // odd ones of its ten halo rows 2q + k, both walking up the columns, so that the accumulator chunks are handed back
    // to the MMA issuer in the order it needs them; a chunk is read as soon as ITS MMAs have retired ----------------
    const int quad = warp & 3, k0 = warp >> 2;
    const int rr = lane >> 4, x = lane & 15, r = 2 * quad + rr;
    const uint32_t tlane = tmem + ((uint32_t)(32 * quad) << 16);
    int tl = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++tl) {
      const int tx = t % a.ntx, tq = t / a.ntx;
      const int ty = tq % a.nty, b = tq / a.nty;
      const int y = ty * TH + r, xg = tx * TW + x;
      const bool pv = y < a.H && xg < a.W;
      float* ob = a.out + (int64_t)b * a.obs + (int64_t)y * a.W + xg;
      int ready = 0, released = 0;
      // the 24 accumulator columns of halo row 2 quad + k; the loads of this warp's next row are in flight while the
      // current one is shifted and stored
      auto issue = [&](float (&v)[24], int k) {
        const int hr = 2 * quad + k;
        while (ready <= (HC * hr + HC - 1) >> 7) {
          const int g = 3 * tl + ready;
          mbar_wait(&tmem_full[g & 3], (g >> 2) & 1);
          ++ready;
          if (threadIdx.x == 0 && ready == 3) CVT_STAMP(tl, 8);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (a.dbg & 2) return;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int c = HC * hr + 8 * j;
          tmem_ld8(tlane + (uint32_t)((((3 * tl + (c >> 7)) & 3) << 7) + (c & 127)), v + 8 * j);
        }
      };
      auto process = [&](const float (&v)[24], int k) {
        const int hr = 2 * quad + k;
        // chunks whose columns lie entirely below this warp's next halo row (hr + 2; its loads are already in flight),
        // and after its last row all of them, are done for this warp
        while (released < 3 && (128 * (released + 1) <= HC * (hr + 2) || k >= 8)) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[(3 * tl + released) & 3]);
          ++released;
        }
        // barrel shift: d[i] = v[x + i], i = 0..8
        float s1[16], s2[12], s3[10], d[9];
#pragma unroll
        for (int i = 0; i < 16; ++i) s1[i] = (x & 8) ? v[i + 8] : v[i];
#pragma unroll
        for (int i = 0; i < 12; ++i) s2[i] = (x & 4) ? s1[i + 4] : s1[i];
#pragma unroll
        for (int i = 0; i < 10; ++i) s3[i] = (x & 2) ? s2[i + 2] : s2[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) d[i] = (x & 1) ? s3[i + 1] : s3[i];
        const int dy = k - rr;                       // halo row hr = r + dy
        if (pv && dy >= 0 && dy <= 8 && !(a.dbg & 1)) {
          const int iy = SGN > 0 ? 8 - dy : dy;      // frame row y - s (iy - 4) = y - 4 + dy
          float* o = ob + (int64_t)iy * hw;
#pragma unroll
          for (int dx = 0; dx < 9; ++dx) {
            const int ix = SGN > 0 ? 8 - dx : dx;
            o[(int64_t)(ix * 9) * hw] = d[dx] * a.kinv;
          }
        }
      };
      float va[24], vb[24];
      issue(va, k0);
#pragma unroll 1
      for (int k = k0; k < k0 + 8; k += 4) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        issue(vb, k + 2);
        process(va, k);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        issue(va, k + 4);
        process(vb, k + 2);
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      process(va, k0 + 8);
      if (threadIdx.x == 0) CVT_STAMP(tl, 9);
    }
  } else if (warp < 14) {
    // ---------------- loaders (warps 12-13) and converters (warps 8-11): operand rows of a 128-row group ----
    // group 0: the reference tile, m = row; groups 1..3: halo positions n = 128 (g - 1) + row
    const bool loader = warp >= 12;
    const int row = threadIdx.x & 127;
    if (loader) {
      // 16-byte cp.async: one copy = four consecutive operand rows (= four consecutive x) of one channel.  Thread =
      // (row quad, channel half): 16 copies per group.  (4-byte copies, one per row and channel, were issue-bound:
      // ~7 cycles per LDGSTS warp instruction, 128 of them per group.)
      const int quad = threadIdx.x & 31, chalf = (threadIdx.x >> 5) & 1;
      int hrow[4], hcol[4];
      hrow[0] = (4 * quad) >> 4; hcol[0] = (4 * quad) & 15;
#pragma unroll
      for (int g = 1; g < 4; ++g) {
        const int n = 128 * (g - 1) + 4 * quad;
        hrow[g] = n / HC - 4;
        hcol[g] = n % HC - 4;
      }
      int item = 0;
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const int tx = t % a.ntx, tq = t / a.ntx;
        const int ty = tq % a.nty, b = tq / a.nty;
        for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
          for (int g = 0; g < 4; ++g, ++item) {
            const int slot = item & 3, use = item >> 2;
            if (use > 0) mbar_wait(&raw_empty[slot], (use - 1) & 1);
            const float* src = (g == 0 ? a.ref : a.frm);
            const int py = ty * TH + hrow[g], px = tx * TW + hcol[g];
            // W % 4 == 0 and px % 4 == 0: the four columns are inside or outside together
            const bool inside = py >= 0 && py < a.H && px >= 0 && px < a.W;
            const int c0 = kc * 32 + chalf * 16;
            const float* p = inside ? src + ((int64_t)b * a.C + c0) * hw + (int64_t)py * a.W + px : src;
            const uint32_t dst = smem_u32(raw + slot * 4096 + (chalf * 16) * 128 + 4 * quad);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const bool ok = inside && c0 + j < a.C;
              const int nbytes = ok ? 16 : 0;
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)j * 512u),
                           "l"(ok ? p + (int64_t)j * hw : src), "r"(nbytes) : "memory");
            }
            cp_async_mbar_arrive_noinc(&raw_full[slot]);
          }
        }
      }
    } else {
      // converter warp g owns operand group g (0: the reference tile A, 1..3: halo chunks B0..B2) and raw slot g: four
      // rows per lane, ONE proxy fence per group and warp.  (fence.proxy.async compiles to MEMBAR.ALL.CTA +
      // FENCE.VIEW.ASYNC: with four warps sharing every group each group paid ~1 000 cycles of fence + barrier latency
      // in sequence; now the four groups of a step are converted side by side.)
      const int g = warp - 8;
      const uint32_t sbase = smem_u32(smem);
      const uint32_t hi0 = sbase + (g == 0 ? OFF_AH : OFF_BH + (g - 1) * 128 * 128);
      const uint32_t lo0 = sbase + (g == 0 ? OFF_AL : OFF_BL + (g - 1) * 128 * 128);
      int step = 0;
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        for (int kc = 0; kc < nkc; ++kc, ++step) {
          mbar_wait(&raw_full[g], step & 1);
          if (lane == 0 && nkc == 1) CVT_STAMP(step, 10 + g);
          uint32_t abo = 0;
          if (g == 0) {
            abo = (uint32_t)((step & 1) * A_STAGE);
            if (step > 1) mbar_wait(&a_empty[step & 1], ((step >> 1) - 1) & 1);
          } else if (step > 0) {
            mbar_wait(&ops_empty[g], (step - 1) & 1);
          }
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int row = lane + 32 * h;
            float v[32];
            const float* src = raw + g * 4096 + row;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = src[j * 128];
            store_row(v, hi0 + abo + (uint32_t)row * 128u, lo0 + abo + (uint32_t)row * 128u, (uint32_t)row & 7u);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&raw_empty[g]);     // release orders the loads above before the slot's refill
            mbar_arrive(g == 0 ? &a_full[step & 1] : &ops_full[g]);
          }
          if (lane == 0 && nkc == 1) CVT_STAMP(step, g);
        }
      }
    }
  } else if (lane == 0) {
    // ---------------- MMA issuer ----------------
    // instruction descriptor: D fp32, A and B TF32, both K-major, N = 128, M = 128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sbase = smem_u32(smem);
    int step = 0, tl = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++tl) {
      for (int kc = 0; kc < nkc; ++kc, ++step) {
        const int ab = step & 1;
        const uint32_t ah = sbase + OFF_AH + (uint32_t)(ab * A_STAGE), al = sbase + OFF_AL + (uint32_t)(ab * A_STAGE);
        mbar_wait(&a_full[ab], (step >> 1) & 1);
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
          mbar_wait(&ops_full[1 + i], step & 1);
          const int g = 3 * tl + i, slot = g & 3, use = g >> 2;
          if (kc == 0 && use > 0) mbar_wait(&tmem_empty[slot], (use - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (kc == 0) CVT_STAMP(tl, 4 + i);
          const uint32_t bh = sbase + OFF_BH + (uint32_t)i * (128u * 128u), bl = sbase + OFF_BL + (uint32_t)i * (128u * 128u);
          const uint32_t dcol = tmem + (uint32_t)(slot * 128);
#pragma unroll
          for (int p = 0; p < 3; ++p) {            // hi * hi, lo * hi, hi * lo
            const uint32_t pa = (p == 1 ? al : ah), pb = (p == 2 ? bl : bh);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_tf32(dcol, smem_desc(pa + 32u * k), smem_desc(pb + 32u * k), idesc, (kc > 0 || p > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&ops_empty[1 + i]);
          if (kc == nkc - 1) umma_commit(&tmem_full[slot]);
        }
        umma_commit(&a_empty[ab]);
      }
      CVT_STAMP(tl, 7);
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace cvt
}  // namespace

// costvol.cu's dispatch calls this for F = 2, win = 9.
int launch_costvol_fwd_tc(const float* ref, const float* frm, float* out, int64_t obs, int B, int C, int H, int W, float kdiv,
                          int sgn, int dbg, cudaStream_t st) {
  cvt::Args a{};
  a.ref = ref; a.frm = frm; a.out = out; a.obs = obs;
  a.C = C; a.H = H; a.W = W;
  a.ntx = (W + cvt::TW - 1) / cvt::TW;
  a.nty = (H + cvt::TH - 1) / cvt::TH;
  const int64_t ntiles = (int64_t)a.ntx * a.nty * B;
  if (ntiles > 0x3fffffff) return fail(B2F_EINVAL, "costvol_forward: too many tiles");
  if ((int64_t)81 * H * W > 0x7fffffff) return fail(B2F_EINVAL, "costvol_forward: image too large for the tensor-core path");
  a.ntiles = (int)ntiles;
  a.kinv = 1.f / kdiv;
  a.trace = tc_trace_buffer();
  a.dbg = dbg;
  static thread_local int attr_dev = -1;
  int dev = 0;
  B2F_CUDA_TRY(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    B2F_CUDA_TRY(cudaFuncSetAttribute(cvt::costvol_fwd_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cvt::SMEM_BYTES));
    B2F_CUDA_TRY(cudaFuncSetAttribute(cvt::costvol_fwd_tc<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cvt::SMEM_BYTES));
    attr_dev = dev;
  }
  const int grid = (int)std::min<int64_t>(ntiles, num_sms());
  if (sgn > 0) cvt::costvol_fwd_tc<1><<<grid, cvt::THREADS, cvt::SMEM_BYTES, st>>>(a);
  else cvt::costvol_fwd_tc<-1><<<grid, cvt::THREADS, cvt::SMEM_BYTES, st>>>(a);
  B2F_CHECK_LAUNCH("costvol_fwd_tc");
  return B2F_OK;
}

}  // namespace b2f
