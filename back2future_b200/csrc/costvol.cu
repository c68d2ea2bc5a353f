// nn.CostVolMulti forward / backward for sm_100a.
//
// Semantics (models/CostVolMulti.lua:49-181): for F input maps, window win (odd),
// n = (win-1)/2, s = +1 (fwd) / -1, channel i = (qx+n)*win + (qy+n)  [x-major]:
//   out[b,i,y,x]       = 1/(C(F-1)) sum_{f=1..F-1} sum_c ref[b,c,y,x] * frame_f[b,c,y-s f qy,x-s f qx]
//   gradRef[b,c,y,x]   = 1/(C(F-1)) sum_f sum_i go[b,i,y,x]           * frame_f[b,c,y-dy,x-dx]
//   gradFrame_f[b,c,p] = 1/(C(F-1))       sum_i go[b,i,p+d]           * ref[b,c,p+d]
// Out-of-range sources are dropped, the normaliser is constant (also at borders).
//
// Two implementations:
//   * tiled TMA kernels for the model's configuration (F = 2, win = 9, W % 4 == 0): feature tiles
//     plus the +-4 halo are staged into shared memory by TMA (out-of-bounds box elements read as
//     zero = "dropped terms"), fp32 FFMA register tiles, no atomics anywhere: the backward is a
//     gather per output element over the 81 displacements (displacement-major slabs of gradOut).
//   * generic direct kernels for every other shape (any F, any odd win, any W).
#include "common.cuh"
#include "tma.cuh"

#include <algorithm>
#include <type_traits>

namespace b2f {
// costvol_tc.cu: the forward on the tensor cores (tcgen05, three-pass TF32 split)
int launch_costvol_fwd_tc(const float* ref, const float* frm, float* out, int64_t obs, int B, int C, int H, int W, float kdiv,
                          int sgn, int dbg, cudaStream_t st);
namespace {

constexpr int kMaxFrames = 8;
struct FramePtrs {
  const float* p[kMaxFrames];
};
struct GradPtrs {
  float* p[kMaxFrames];
};

// =======================================================================================
// generic kernels
// =======================================================================================

__global__ void __launch_bounds__(256)
costvol_fwd_generic(FramePtrs fr, int F, int B, int C, int H, int W, int win, int sgn,
                    float* __restrict__ out, int64_t obs, float kdiv) {
  const int n = (win - 1) / 2;
  const int win2 = win * win;
  const int64_t hw = (int64_t)H * W;
  const int64_t total = (int64_t)B * win2 * hw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int i = (int)((idx / hw) % win2);
    const int b = (int)(idx / (hw * win2));
    const int qx = i / win - n, qy = i % win - n;
    const float* r = fr.p[0] + (int64_t)b * C * hw + (int64_t)y * W + x;
    float acc = 0.f;
    for (int f = 1; f < F; ++f) {
      const int xs = x - sgn * f * qx, ys = y - sgn * f * qy;
      if (xs < 0 || xs >= W || ys < 0 || ys >= H) continue;
      const float* g = fr.p[f] + (int64_t)b * C * hw + (int64_t)ys * W + xs;
      for (int c = 0; c < C; ++c) acc = fmaf(__ldg(r + c * hw), __ldg(g + c * hw), acc);
    }
    out[(int64_t)b * obs + (int64_t)i * hw + (int64_t)y * W + x] = acc / kdiv;
  }
}

// which == 0: gradient of the reference map; which == f >= 1: gradient of frame f.
__global__ void __launch_bounds__(256)
costvol_bwd_generic(FramePtrs fr, int F, int which, int B, int C, int H, int W, int win, int sgn,
                    const float* __restrict__ go, int64_t gbs, float* __restrict__ gout, float kdiv) {
  const int n = (win - 1) / 2;
  const int64_t hw = (int64_t)H * W;
  const int64_t total = (int64_t)B * C * hw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int c = (int)((idx / hw) % C);
    const int b = (int)(idx / (hw * C));
    const float* gob = go + (int64_t)b * gbs;
    float acc = 0.f;
    if (which == 0) {
      for (int f = 1; f < F; ++f) {
        const float* g = fr.p[f] + ((int64_t)b * C + c) * hw;
        for (int ix = 0; ix < win; ++ix) {
          const int xs = x - sgn * f * (ix - n);
          if (xs < 0 || xs >= W) continue;
          for (int iy = 0; iy < win; ++iy) {
            const int ys = y - sgn * f * (iy - n);
            if (ys < 0 || ys >= H) continue;
            acc = fmaf(__ldg(gob + (int64_t)(ix * win + iy) * hw + (int64_t)y * W + x),
                       __ldg(g + (int64_t)ys * W + xs), acc);
          }
        }
      }
    } else {
      const int f = which;
      const float* r = fr.p[0] + ((int64_t)b * C + c) * hw;
      for (int ix = 0; ix < win; ++ix) {
        const int xq = x + sgn * f * (ix - n);
        if (xq < 0 || xq >= W) continue;
        for (int iy = 0; iy < win; ++iy) {
          const int yq = y + sgn * f * (iy - n);
          if (yq < 0 || yq >= H) continue;
          const int64_t o = (int64_t)yq * W + xq;
          acc = fmaf(__ldg(gob + (int64_t)(ix * win + iy) * hw + o), __ldg(r + o), acc);
        }
      }
    }
    gout[idx] = acc / kdiv;
  }
}

// =======================================================================================
// tiled TMA forward (F = 2, win = 9)
// =======================================================================================
//
// CTA = one 8 x (4A) pixel tile of one batch item; 9 compute warps (one per dy) + 1 TMA producer
// warp.  Lane = (row r = lane & 7, strip st = lane >> 3); a thread owns A consecutive pixels of
// row r and the 9 dx displacements of its warp's dy: 9*A accumulators.  Channels stream through a
// 3-stage full/empty mbarrier ring in chunks of 8: per channel a thread reads A ref values and
// A+8 frame values (LDS.128) for 9*A FFMAs.  Box row pitches are (odd * 16) bytes so the eight
// rows of a quarter-warp hit disjoint bank groups.
namespace cvf {
constexpr int TH = 8;    // tile rows
constexpr int CK = 8;    // channels per pipeline stage
constexpr int NCW = 9;   // compute warps = window rows
constexpr int THREADS = (NCW + 1) * 32;

// PIPE = true : one CTA per SM, up to 168 registers, operand registers double-buffered (the loads of
//               channel c+1 are in flight under the FFMAs of channel c), 6-stage ring.
// PIPE = false: two CTAs per SM, 96 registers, 3-stage ring (small levels: more CTAs, latency-bound anyway).
template <int A, bool PIPE>
struct Cfg {
  static constexpr int NS = PIPE ? 4 : 3;   // pipeline stages
  static constexpr int TW = 4 * A;
  static constexpr int RW = TW + 4;         // ref box width   (RW/4 odd)
  static constexpr int FW = TW + 12;        // frame box width (FW/4 odd), covers the +-4 halo
  static constexpr int FR = TH + 8;         // frame box rows
  static constexpr int REF_ELEMS = CK * TH * RW;
  static constexpr int FRM_ELEMS = CK * FR * FW;
  static constexpr int STAGE_ELEMS = REF_ELEMS + FRM_ELEMS;
  static constexpr int STAGE_BYTES = STAGE_ELEMS * 4;
  // PIPE: the 81 x 8 x 32 output tile is staged in shared memory (128-byte swizzle) and written by one TMA
  // store, so the compute warps go straight on to the next tile
  static constexpr int OUT_BYTES = PIPE ? 81 * TH * TW * 4 : 0;
  static constexpr int SMEM_BYTES = OUT_BYTES + NS * STAGE_BYTES + 2 * NS * 8 + 1024;
  static constexpr int CTAS_PER_SM = PIPE ? 1 : 2;
  // register budget: the register file is 16K per scheduler and warps are placed four at a time, so
  // 2 CTAs x 10 warps -> 5 warps per scheduler x 96; 1 CTA x 10 warps -> 3 per scheduler x 168
  static constexpr int MAXREG = PIPE ? 168 : 96;
  static_assert((RW / 4) % 2 == 1 && (FW / 4) % 2 == 1, "row pitch must be odd*16B");
  static_assert((REF_ELEMS * 4) % 128 == 0 && (FRM_ELEMS * 4) % 128 == 0, "TMA dst alignment");
  static_assert(SMEM_BYTES * CTAS_PER_SM <= 227 * 1024, "shared memory budget");
  static_assert(!PIPE || TW * 4 == 128, "the swizzled staging tile needs 128-byte rows");
  static_assert(OUT_BYTES % 1024 == 0 && STAGE_BYTES % 128 == 0, "alignment");
};

template <int A>
struct Operands {
  float rv[A], fv[A + 8];
};

template <int A>
__device__ __forceinline__ void load_operands(Operands<A>& o, uint32_t rs, uint32_t fs) {
#pragma unroll
  for (int q = 0; q < A / 4; ++q) {
    const float4 v = lds128(rs + 16u * q);
    o.rv[4 * q] = v.x; o.rv[4 * q + 1] = v.y; o.rv[4 * q + 2] = v.z; o.rv[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int q = 0; q < (A + 8) / 4; ++q) {
    const float4 v = lds128(fs + 16u * q);
    o.fv[4 * q] = v.x; o.fv[4 * q + 1] = v.y; o.fv[4 * q + 2] = v.z; o.fv[4 * q + 3] = v.w;
  }
}

template <int A, int SGN>
__device__ __forceinline__ void fma_window_row(float (&acc)[9][A], const Operands<A>& o) {
#pragma unroll
  for (int ix = 0; ix < 9; ++ix)
#pragma unroll
    for (int j = 0; j < A; ++j)
      acc[ix][j] = fmaf(o.rv[j], o.fv[j + (SGN > 0 ? 8 - ix : ix)], acc[ix][j]);
}

// FFMA2 form (PIPE kernels).  For pixel j the nine window columns read the nine consecutive frame values
// fv[j + d], d = 0..8 (d = ix when SGN < 0, 8 - ix when SGN > 0), all multiplied by the same ref value rv[j].
// Two neighbouring d form one FFMA2 whose multiplicand pair is a LOADED, even-aligned frame pair and whose
// multiplier is rv[j] broadcast (the scalar-operand form of FFMA2): d pairs (0,1)(2,3)(4,5)(6,7) + scalar d = 8
// for even j, scalar d = 0 + pairs (1,2)(3,4)(5,6)(7,8) for odd j.  No register shuffling at all; per channel
// 32 FFMA2 + 8 FFMA instead of 72 FFMA: about half the issue slots and register-file reads per FMA (measured on
// B200: 102 vs 87 FMA/clk/SM, tools/ubench/smem_fma.cu).  Each half is an fp32 fma, rounded exactly as fmaf.
template <int A>
struct Operands2 {
  float rv[A];
  f32x2 fe[(A + 8) / 2];   // fe[m] = (fv[2m], fv[2m+1])
};

template <int A>
__device__ __forceinline__ void load_operands2(Operands2<A>& o, uint32_t rs, uint32_t fs) {
#pragma unroll
  for (int q = 0; q < A / 4; ++q) {
    const float4 v = lds128(rs + 16u * q);
    o.rv[4 * q] = v.x; o.rv[4 * q + 1] = v.y; o.rv[4 * q + 2] = v.z; o.rv[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int q = 0; q < (A + 8) / 4; ++q) lds128_pairs(fs + 16u * q, o.fe[2 * q], o.fe[2 * q + 1]);
}

// accumulators of one thread: A pixels x 9 frame offsets d
template <int A>
struct Acc2 {
  f32x2 pair[A][4];   // pixel j, offsets (2m + (j&1), 2m + 1 + (j&1))
  float single[A];    // pixel j, offset 8 (j even) or 0 (j odd)
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int j = 0; j < A; ++j) {
#pragma unroll
      for (int m = 0; m < 4; ++m) pair[j][m] = 0ull;
      single[j] = 0.f;
    }
  }
  // value for pixel j at frame offset d (compile-time indices only)
  __device__ __forceinline__ float get(int j, int d) const {
    const int odd = j & 1;
    if (d == (odd ? 0 : 8)) return single[j];
    float lo, hi;
    unpack2(pair[j][(d - odd) >> 1], lo, hi);
    return ((d - odd) & 1) ? hi : lo;
  }
};

template <int A>
__device__ __forceinline__ void fma2_window_row(Acc2<A>& acc, const Operands2<A>& o) {
#pragma unroll
  for (int j = 0; j < A; ++j) {
    const int odd = j & 1;
    const f32x2 s = pack2(o.rv[j], o.rv[j]);
#pragma unroll
    for (int m = 0; m < 4; ++m) fma2_acc(acc.pair[j][m], s, o.fe[(j + odd + 2 * m) >> 1]);
    float lo, hi;
    unpack2(o.fe[odd ? (j - 1) >> 1 : (j + 8) >> 1], lo, hi);
    acc.single[j] = fmaf(o.rv[j], odd ? hi : lo, acc.single[j]);
  }
}

// Persistent kernel: CTA c takes tiles c, c + gridDim.x, ... in raster order (x fastest), so the CTAs active
// at any moment sit on neighbouring tiles and share their halos in L2.  The channel chunks of a CTA's
// consecutive tiles flow through ONE ring: the producer is already fetching the next tile while the consumers
// finish (and store) the current one.
// Small levels: the channel range is split over nsplit work items whose partial sums are reduced with
// red.global.add into a pre-zeroed output (enough items to fill 148 SMs from a handful of tiles).
template <int A, int SGN, bool PIPE>
__global__ void __maxnreg__((Cfg<A, PIPE>::MAXREG))
costvol_fwd_tma(const __grid_constant__ CUtensorMap tm_ref, const __grid_constant__ CUtensorMap tm_frm,
                const __grid_constant__ CUtensorMap tm_out, float* __restrict__ out, int64_t obs, int C, int H, int W, float kdiv, int nsplit,
                int chunks_per_split, int ntx, int nty, int ntiles, int dbg) {
  using cfg = Cfg<A, PIPE>;
  constexpr int NS = cfg::NS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // 32-bit shared addresses throughout (a generic pointer costs an S2R + LEA per access in the loop)
  const uint32_t obase = smem_u32(smem);               // output staging tile (PIPE), 1024-byte aligned
  const uint32_t sbase = obase + cfg::OUT_BYTES;       // ring stages
  const uint32_t full0 = sbase + NS * cfg::STAGE_BYTES, empty0 = full0 + 8u * NS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks_all = (C + CK - 1) / CK;

  if (threadIdx.x == 0) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cfg::OUT_BYTES + NS * cfg::STAGE_BYTES);
    for (int i = 0; i < NS; ++i) {
      mbar_init(bars + i, 1);
      mbar_init(bars + NS + i, NCW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == NCW) {  // ---- TMA producer ----
    if (lane == 0) {
      tma_prefetch_desc(&tm_ref);
      tma_prefetch_desc(&tm_frm);
      int s = 0;
      uint32_t ph = 0;
      bool wrapped = false;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = t % ntx, tq = t / ntx;
        const int ty = tq % nty, bz = tq / nty;
        const int b = bz / nsplit, split = bz - b * nsplit;
        const int x0 = tx * cfg::TW, y0 = ty * TH;
        const int kbeg = split * chunks_per_split;
        const int nchunks = min(nchunks_all - kbeg, chunks_per_split);
        for (int k = 0; k < nchunks; ++k) {
          if (wrapped) mbar_wait_addr(empty0 + 8u * s, ph);
          const uint32_t rs = sbase + (uint32_t)s * cfg::STAGE_BYTES;
          const uint32_t fs = rs + 4u * cfg::REF_ELEMS;
          mbar_expect_tx_addr(full0 + 8u * s, cfg::STAGE_BYTES);
          tma_load_4d_addr(rs, &tm_ref, x0, y0, (kbeg + k) * CK, b, full0 + 8u * s);
          tma_load_4d_addr(fs, &tm_frm, x0 - 4, y0 - 4, (kbeg + k) * CK, b, full0 + 8u * s);
          if (++s == NS) {
            s = 0;
            if (wrapped) ph ^= 1u;
            wrapped = true;
          }
        }
      }
    }
    return;
  }

  // ---- consumers ----
  const int iy = warp;
  const int r = lane & 7, st = lane >> 3;
  const int frow = r + 4 - SGN * (iy - 4);  // frame box row of this thread's source pixels
  // THC's div(scalar) multiplies floats by the reciprocal; so do we.
  const float kinv = 1.f / kdiv;
  const int64_t hw = (int64_t)H * W;
  const uint32_t roff = sbase + 4u * (r * cfg::RW + A * st);
  const uint32_t foff = sbase + 4u * (cfg::REF_ELEMS + frow * cfg::FW + A * st);
  constexpr uint32_t RSTEP = 4u * (TH * cfg::RW), FSTEP = 4u * (cfg::FR * cfg::FW);

  int s = 0;            // ring position of the next chunk
  uint32_t ph = 0;      // its phase parity
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tx = t % ntx, tq = t / ntx;
    const int ty = tq % nty, bz = tq / nty;
    const int b = bz / nsplit, split = bz - b * nsplit;
    const int nchunks = min(nchunks_all - split * chunks_per_split, chunks_per_split);

    // PIPE: packed accumulators (FFMA2 form); else scalar (FFMA form)
    Acc2<PIPE ? A : 4> acc2;
    float acc[PIPE ? 1 : 9][PIPE ? 1 : A];
    if constexpr (PIPE) {
      acc2.clear();
    } else {
#pragma unroll
      for (int ix = 0; ix < 9; ++ix)
#pragma unroll
        for (int j = 0; j < A; ++j) acc[ix][j] = 0.f;
    }

    for (int k = 0; k < nchunks; ++k) {
      mbar_wait_addr(full0 + 8u * s, ph);
      uint32_t rs = roff + (uint32_t)s * cfg::STAGE_BYTES;
      uint32_t fs = foff + (uint32_t)s * cfg::STAGE_BYTES;
      if (dbg & 1) {
        // measurement aid (b2f_debug_costvol_path 8/10): TMA feed only, no arithmetic
      } else if constexpr (PIPE) {
        // two operand sets: the loads of the next channel are issued before the FFMAs of the current one
        Operands2<A> o0, o1;
        load_operands2<A>(o0, rs, fs);
        // fully unrolled: with one trip per two channels ptxas closed every trip with ~30 MOVs (phi copies of
        // accumulator pairs it had renamed inside the body: 19 % of the loop's issue slots)
#pragma unroll
        for (int cc = 0; cc < CK; cc += 2) {
          load_operands2<A>(o1, rs + RSTEP, fs + FSTEP);
          fma2_window_row<A>(acc2, o0);
          rs += 2 * RSTEP;
          fs += 2 * FSTEP;
          if (cc + 2 < CK) load_operands2<A>(o0, rs, fs);
          fma2_window_row<A>(acc2, o1);
        }
      } else {
#pragma unroll 1
        for (int cc = 0; cc < CK; ++cc) {
          Operands<A> o;
          load_operands<A>(o, rs, fs);
          rs += RSTEP;
          fs += FSTEP;
          fma_window_row<A, SGN>(acc, o);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_addr(empty0 + 8u * s);
      if (++s == NS) {
        s = 0;
        ph ^= 1u;
      }
    }

    // ---- epilogue: out[b, ix*9+iy, y, x0 + A*st + j] ----
    if constexpr (PIPE) {
      // stage the tile as [channel][row][32 columns] with the TMA 128-byte swizzle (16-byte chunk c of row R
      // sits at chunk c ^ (R & 7)): the eight rows of a quarter-warp hit eight different bank groups
#ifndef B2F_CVF_WARP_STORE
#define B2F_CVF_WARP_STORE 1
#endif
      if (B2F_CVF_WARP_STORE && (dbg & 4)) {
        // Per-warp epilogue (tm_out is then the 5-D view (x, y, window row iy, window column ix, batch)): warp iy
        // stages its nine planes ix*9+iy in its own [ix][row][128 B] region and issues its own bulk tensor store, so
        // no warp waits for another one -- with 9 compute warps on 4 schedulers (3/2/2/2) the two CTA-wide barriers
        // of the whole-tile store were the hottest spot of the kernel (15 % of the stall samples on the first STS
        // behind the barrier).  Bulk async groups are per thread: lane 0 waits for ITS previous store only.
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        const uint32_t wbase = obase + (uint32_t)(iy * 9 * TH * 128);
#pragma unroll
        for (int ix = 0; ix < 9; ++ix) {
          const uint32_t rowaddr = wbase + (uint32_t)((ix * TH + r) * 128);
#pragma unroll
          for (int q = 0; q < A / 4; ++q) {
            const int d = SGN > 0 ? 8 - ix : ix;
            const float4 v = make_float4(acc2.get(4 * q, d) * kinv, acc2.get(4 * q + 1, d) * kinv,
                                         acc2.get(4 * q + 2, d) * kinv, acc2.get(4 * q + 3, d) * kinv);
            sts128(rowaddr + 16u * (uint32_t)(((A / 4) * st + q) ^ r), v);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(dbg & 2)) {
          if (nsplit == 1) tma_store_5d_addr(wbase, &tm_out, tx * cfg::TW, ty * TH, iy, 0, b);
          else tma_reduce_add_5d_addr(wbase, &tm_out, tx * cfg::TW, ty * TH, iy, 0, b);
          tma_store_commit();
        }
      } else {
      if (warp == 0) {   // the previous tile's TMA store must have finished reading the staging tile
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      named_bar_sync(1, NCW * 32);
#pragma unroll
      for (int ix = 0; ix < 9; ++ix) {
        const uint32_t rowaddr = obase + (uint32_t)(((ix * 9 + iy) * TH + r) * 128);
#pragma unroll
        for (int q = 0; q < A / 4; ++q) {
          const int d = SGN > 0 ? 8 - ix : ix;
          const float4 v = make_float4(acc2.get(4 * q, d) * kinv, acc2.get(4 * q + 1, d) * kinv,
                                       acc2.get(4 * q + 2, d) * kinv, acc2.get(4 * q + 3, d) * kinv);
          sts128(rowaddr + 16u * (uint32_t)(((A / 4) * st + q) ^ r), v);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, NCW * 32);
      if (warp == 0 && lane == 0 && !(dbg & 2)) {
        if (nsplit == 1) tma_store_4d_addr(obase, &tm_out, tx * cfg::TW, ty * TH, 0, b);
        else tma_reduce_add_4d_addr(obase, &tm_out, tx * cfg::TW, ty * TH, 0, b);
        tma_store_commit();
      }
      }
    } else {
      const int y = ty * TH + r;
      if (y < H && !(dbg & 2)) {
        const int xb = tx * cfg::TW + A * st;
        float* ob = out + (int64_t)b * obs + (int64_t)y * W + xb + (int64_t)iy * hw;
#pragma unroll
        for (int ix = 0; ix < 9; ++ix) {
          float* o = ob + (int64_t)(ix * 9) * hw;
#pragma unroll
          for (int q = 0; q < A / 4; ++q) {
            if (xb + 4 * q < W) {
              const float4 v = make_float4(acc[ix][4 * q] * kinv, acc[ix][4 * q + 1] * kinv,
                                           acc[ix][4 * q + 2] * kinv, acc[ix][4 * q + 3] * kinv);
              if (nsplit == 1) {
                *reinterpret_cast<float4*>(o + 4 * q) = v;
              } else {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * q), "f"(v.x), "f"(v.y),
                             "f"(v.z), "f"(v.w)
                             : "memory");
              }
            }
          }
        }
      }
    }
  }
  if (PIPE && lane == 0) tma_store_wait_read();   // shared memory must outlive the last store (of every warp)
}
}  // namespace cvf

// =======================================================================================
// tiled TMA backward (F = 2, win = 9)
// =======================================================================================
//
// out[c,p] = 1/k sum_q G_q[p] * X[c, p + T q]   with
//   role 0 (gradRef):   X = frame, T = -s, G_q[p] = go[q, p]
//   role 1 (gradFrame): X = ref,   T = +s, G_q[p] = go[q, p + s q]   (the shift is the TMA box origin)
// CTA = (8 x 32 pixel tile, 32-channel chunk, role); 4 compute warps (8 channels each) + 1 TMA
// producer warp.  A thread owns 8 pixels x 8 channels = 64 accumulators.  The X halo tile of the
// chunk stays resident; gradOut arrives as one slab per window row iy (9 boxes of 8 x 36, double
// buffered).  Per slab a thread reads its 9 x 8 gradOut values once and, per channel, 16 X values
// for 72 FFMAs.  Every output element is written exactly once: no atomics, no zero-fill.
namespace cvb {
constexpr int TH = 8;             // tile rows
constexpr int CC = 32;            // channels per CTA
constexpr int CG = 8;             // channels per warp
constexpr int NCG = CC / CG;      // channel groups
constexpr int XR = TH + 8;        // X box rows
__host__ __device__ constexpr int fdiv4(int v) { return (v >= 0) ? v / 4 : -((3 - v) / 4); }

// TW = tile width, NSLAB = gradOut slab ring depth:
//   <32, 2>: two CTAs per SM (115.6 KB each: exactly the limit), four compute warps -- narrow levels
//   <32, 9>: one CTA per SM, all nine slabs requested at once -- small levels (a chain of round trips)
//   <64, 3>: one CTA per SM, eight compute warps (two 32-column halves x four channel groups), 3-deep ring.
//            The TMA unit delivers one box ROW per ~8 cycles per SM whatever its length
//            (tools/ubench/tma_feed.cu), so 64-column tiles halve the feed time per pixel -- but the kernel
//            as a whole measured slower than <32, 2> (see the dispatch), so it is an experiment, not the default
template <int TW_, int NSLAB_>
struct Cfg {
  static constexpr int TW = TW_, NSLAB = NSLAB_;
  static constexpr int NH = TW / 32;              // 32-column halves
  static constexpr int NCW = NCG * NH;            // compute warps
  static constexpr int THREADS = (NCW + 1) * 32;
  static constexpr int XW = TW + 12;              // X box width (odd multiple of 16 B)
  static constexpr int GW0 = TW + 4;              // gradRef:   gradOut box at (x0, y0)
  static constexpr int GW1 = TW + 12;             // gradFrame: aligned superset at (x0-4, y0+s*qy)
  static constexpr int XG_ELEMS = CG * XR * XW;   // one channel group of X
  static constexpr int X_ELEMS = NCG * XG_ELEMS;
  static constexpr int GBOX_ELEMS = TH * GW1;     // slab geometry is sized for the wider role
  static constexpr int SLAB_ELEMS = 9 * GBOX_ELEMS;
  static constexpr int XH_ELEMS = CG * (XR / 2) * XW;   // one half (8 rows) of a channel group: its own TMA box + mbarrier
  static constexpr int NBAR = 2 * NCG + 2 * NSLAB;
  static constexpr int SMEM_BYTES = (X_ELEMS + NSLAB * SLAB_ELEMS) * 4 + NBAR * 8 + 128;
  static constexpr int CTAS_PER_SM = (TW == 32 && NSLAB == 2) ? 2 : 1;
  static_assert((XW / 4) % 2 == 1 && (GW0 / 4) % 2 == 1 && (GW1 / 4) % 2 == 1, "row pitch must be odd*16B");
  static_assert((XG_ELEMS * 4) % 128 == 0 && (XH_ELEMS * 4) % 128 == 0 && (GBOX_ELEMS * 4) % 128 == 0, "TMA dst alignment");
  static_assert(SMEM_BYTES * CTAS_PER_SM <= 228 * 1024 - 1024 * CTAS_PER_SM && SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// Copy this thread's 9 x 8 gradOut values of the current slab into registers.  SHIFT == 0 (gradRef): the
// slab holds go[q] at the tile itself.  SHIFT == +-1 (gradFrame): the slab holds the 16-byte aligned
// superset [x0-4, x0+TW+8) of the rows shifted by s*qy; the x shift s*qx is a compile-time constant per ix,
// so the 8 wanted floats are 2-3 aligned LDS.128 plus a static register selection (a TMA box cannot start
// at an x that is not a multiple of 16 bytes).  `col` = first column of the thread inside the tile.
template <class cfg, int SHIFT>
__device__ __forceinline__ void slab_load_g(float (&g)[9][8], const float* __restrict__ slab, int r, int col) {
  constexpr int GWS = SHIFT == 0 ? cfg::GW0 : cfg::GW1;
#pragma unroll
  for (int ix = 0; ix < 9; ++ix) {
    const int base = SHIFT == 0 ? 0 : 4 + SHIFT * (ix - 4);   // offset of the window inside the box row
    const int o = base - 4 * fdiv4(base);
    const float* p = slab + ix * (TH * GWS) + r * GWS + col + 4 * fdiv4(base);
    float v[12];
#pragma unroll
    for (int q = 0; q < (o == 0 ? 2 : 3); ++q) {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[ix][j] = v[j + o];
  }
}

template <class cfg>
__device__ __forceinline__ void load_xv(float (&xv)[16], const float* __restrict__ xs, int c) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(xs + c * ((XR / 2) * cfg::XW) + 4 * q);
    xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
  }
}
#ifndef B2F_CVB_DB
#define B2F_CVB_DB 1
#endif
// The X operands of channel c + 1 are requested before the 72 FMAs of channel c (B2F_CVB_DB, default on): with two
// compute warps per scheduler an LDS round trip in front of every channel's FMAs is not always covered.
template <class cfg, int T>
__device__ __forceinline__ void slab_fma(float (&acc)[CG][8], const float (&g)[9][8], const float* __restrict__ xs) {
#if B2F_CVB_DB
  float xa[16], xb[16];
  load_xv<cfg>(xa, xs, 0);
#pragma unroll
  for (int c = 0; c < CG; c += 2) {
    load_xv<cfg>(xb, xs, c + 1);
#pragma unroll
    for (int ix = 0; ix < 9; ++ix)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        acc[c][j] = fmaf(g[ix][j], xa[j + (T > 0 ? ix : 8 - ix)], acc[c][j]);
    if (c + 2 < CG) load_xv<cfg>(xa, xs, c + 2);
#pragma unroll
    for (int ix = 0; ix < 9; ++ix)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        acc[c + 1][j] = fmaf(g[ix][j], xb[j + (T > 0 ? ix : 8 - ix)], acc[c + 1][j]);
  }
#else
#pragma unroll
  for (int c = 0; c < CG; ++c) {
    float xv[16];
    load_xv<cfg>(xv, xs, c);
#pragma unroll
    for (int ix = 0; ix < 9; ++ix)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        acc[c][j] = fmaf(g[ix][j], xv[j + (T > 0 ? ix : 8 - ix)], acc[c][j]);
  }
#endif
}

// Timeline instrumentation (tools/cvb_trace.py; only in builds with -DB2F_CVB_TRACE, tools/build_variants.sh): lane 0
// of every warp stores clock64 stamps -- 32 slots per warp, 8 warps per CTA -- into the buffer registered with
// b2f_debug_cvb_trace.  clock64 is per SM: stamps of CTAs with the same %smid (slot 1) share a time base.
#ifdef B2F_CVB_TRACE
__device__ unsigned long long* g_cvb_trace;
#define CVB_STAMP(k)                                                                                         \
  do {                                                                                                       \
    if (lane == 0 && g_cvb_trace)                                                                            \
      g_cvb_trace[((size_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 8 + warp) * 32 + (k)] = clock64(); \
  } while (0)
#define CVB_STAMP_VAL(k, v)                                                                                  \
  do {                                                                                                       \
    if (lane == 0 && g_cvb_trace)                                                                            \
      g_cvb_trace[((size_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 8 + warp) * 32 + (k)] = (v); \
  } while (0)
#else
#define CVB_STAMP(k) do {} while (0)
#define CVB_STAMP_VAL(k, v) do {} while (0)
#endif

template <int SGN, int TW, int NSLAB>
__global__ void __launch_bounds__((Cfg<TW, NSLAB>::THREADS), (Cfg<TW, NSLAB>::CTAS_PER_SM))
costvol_bwd_tma(const __grid_constant__ CUtensorMap tm_frame, const __grid_constant__ CUtensorMap tm_ref,
                const __grid_constant__ CUtensorMap tm_go, const __grid_constant__ CUtensorMap tm_go1,
                const __grid_constant__ CUtensorMap tm_gref, const __grid_constant__ CUtensorMap tm_gfrm,
                int nroles, int role0, int nchunk, int C, int H, int W, float kdiv, int dbg) {
  using cfg = Cfg<TW, NSLAB>;
  constexpr int NCW = cfg::NCW, XW = cfg::XW, GW0 = cfg::GW0, GW1 = cfg::GW1;
  constexpr int XG_ELEMS = cfg::XG_ELEMS, SLAB_ELEMS = cfg::SLAB_ELEMS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // pointer + integer offset keeps the shared address space (LDS, not generic LD)
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float* xsm = reinterpret_cast<float*>(smem);
  float* gsm = xsm + cfg::X_ELEMS;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(smem + (cfg::X_ELEMS + NSLAB * SLAB_ELEMS) * 4);   // [channel group]
  uint64_t* gfull = xfull + 2 * NCG;   // xfull[2 * group + half]
  uint64_t* gempty = gfull + NSLAB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  int z = blockIdx.z;
  const int role = (nroles == 2) ? (z & 1) : role0;
  if (nroles == 2) z >>= 1;
  const int chunk = z % nchunk, b = z / nchunk;
  const int c0 = chunk * CC;
  const int T = (role == 0) ? -SGN : SGN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NCG; ++i) mbar_init(&xfull[i], 1);
    for (int i = 0; i < NSLAB; ++i) {
      mbar_init(&gfull[i], 1);
      mbar_init(&gempty[i], NCW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  CVB_STAMP(0);
#ifdef B2F_CVB_TRACE
  {
    unsigned sm_, gt_lo, gt_hi;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));
    unsigned long long gt_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));
    (void)gt_lo; (void)gt_hi;
    CVB_STAMP_VAL(1, (unsigned long long)sm_);
    CVB_STAMP_VAL(30, gt_);
  }
#endif
  if (warp == NCW) {  // ---- TMA producer ----
    // NOTE: tensor maps are addressed as kernel parameters at every use (a runtime-selected descriptor
    // pointer is not a constant-bank address any more and the TMA faults).
    if (lane == 0) {
      auto load_slab = [&](int iy) {
        const int s = iy % NSLAB;
        float* dst = gsm + s * SLAB_ELEMS;
        if (role == 0) {
          mbar_arrive_expect_tx(&gfull[s], 9 * TH * GW0 * 4);
#pragma unroll 1
          for (int ix = 0; ix < 9; ++ix)
            tma_load_4d(dst + ix * (TH * GW0), &tm_go, x0, y0, ix * 9 + iy, b, &gfull[s]);
        } else {
          mbar_arrive_expect_tx(&gfull[s], 9 * TH * GW1 * 4);
#pragma unroll 1
          for (int ix = 0; ix < 9; ++ix)
            tma_load_4d(dst + ix * (TH * GW1), &tm_go1, x0 - 4, y0 + SGN * (iy - 4), ix * 9 + iy, b, &gfull[s]);
        }
      };
      load_slab(0);
      CVB_STAMP(2);
      // X halo tile, per channel group as two 8-row boxes with their own barriers, the half that window row 0 reads
      // first for every group: slab 0 touches rows 0..7 (T > 0) or 8..15 (T < 0) only, so a warp starts after a
      // quarter of the box rows it used to wait for (the wait for the whole 16-row box was the hottest instruction
      // of the kernel: 10.5 % of the stall samples)
      const int h0 = T > 0 ? 0 : 1;
#pragma unroll 1
      for (int k = 0; k < 2; ++k) {
        const int hf = k == 0 ? h0 : 1 - h0;
        for (int w = 0; w < NCG; ++w) {
          mbar_arrive_expect_tx(&xfull[2 * w + hf], cfg::XH_ELEMS * 4);
          float* dst = xsm + w * XG_ELEMS + hf * cfg::XH_ELEMS;
          if (role == 0) tma_load_4d(dst, &tm_frame, x0 - 4, y0 - 4 + 8 * hf, c0 + w * CG, b, &xfull[2 * w + hf]);
          else           tma_load_4d(dst, &tm_ref, x0 - 4, y0 - 4 + 8 * hf, c0 + w * CG, b, &xfull[2 * w + hf]);
        }
      }
      CVB_STAMP(20);
      for (int iy = 1; iy < 9; ++iy) {
        if (iy >= NSLAB) mbar_wait(&gempty[iy % NSLAB], ((iy / NSLAB) - 1) & 1);
        load_slab(iy);
        CVB_STAMP(2 + iy);
      }
    }
    return;
  }

  // ---- consumers: warp = (channel group, 32-column half) ----
  const int cg = warp % NCG, half = warp / NCG;
  const int r = lane & 7, st = lane >> 3;
  const int col = 32 * half + 8 * st;      // first column of this thread inside the tile
  float acc[CG][8];
#pragma unroll
  for (int c = 0; c < CG; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;

  const float* xbase = xsm + cg * XG_ELEMS + col;
#pragma unroll 1
  for (int iy = 0; iy < 9; ++iy) {
    const int s = iy % NSLAB;
    if (iy == 0) mbar_wait(&xfull[2 * cg + (T > 0 ? 0 : 1)], 0);
    if (iy == 1) mbar_wait(&xfull[2 * cg + (T > 0 ? 1 : 0)], 0);
    mbar_wait(&gfull[s], (iy / NSLAB) & 1);
    CVB_STAMP(2 + 2 * iy);
    const int xrow = r + 4 + T * (iy - 4);     // 0..15: half = row / 8, layout [half][channel][8 rows][XW]
    const float* xs = xbase + (xrow >> 3) * cfg::XH_ELEMS + (xrow & 7) * XW;
    float g[9][8];
    if (role == 0) slab_load_g<cfg, 0>(g, gsm + s * SLAB_ELEMS, r, col);
    else           slab_load_g<cfg, SGN>(g, gsm + s * SLAB_ELEMS, r, col);
    // the slab is dead as soon as every lane holds its values: release it BEFORE the FFMAs so the
    // producer refills it a whole window row earlier
    __syncwarp();
    if (lane == 0) mbar_arrive(&gempty[s]);
    if (dbg & 1) continue;   // measurement aid: feed only
    if (T > 0) slab_fma<cfg, 1>(acc, g, xs);
    else       slab_fma<cfg, -1>(acc, g, xs);
    CVB_STAMP(3 + 2 * iy);
  }

  // ---- epilogue ----
  // The X group is dead once every warp that reads it is done (with 64-column tiles the two halves of a channel
  // group share it: a 64-thread named barrier).  A warp's 8 channels x 8 rows x 32 columns of results are staged
  // in its part of the group with the TMA 128-byte swizzle (16-byte chunk c of the 128-byte row at address A
  // sits at chunk c ^ ((A >> 7) & 7), so the eight rows of a quarter-warp hit eight different bank groups) and
  // leave as ONE bulk tensor store; channels >= C and pixels outside the image are clipped by the TMA.  (Direct
  // STG.128 from this accumulator layout touches 8 lines per instruction: 33 cycles each, 16 us of the 117 us
  // kernel at level 3.)
  if (cfg::NH > 1) named_bar_sync(1 + cg, 32 * cfg::NH);
  const float kinv = 1.f / kdiv;
  const uint32_t wbase = smem_u32(xsm) + (uint32_t)(cg * XG_ELEMS) * 4u + (uint32_t)half * (CG * TH * 128);
#pragma unroll
  for (int c = 0; c < CG; ++c) {
    const uint32_t rowaddr = wbase + (uint32_t)((c * TH + r) * 128);
    const uint32_t sw = (rowaddr >> 7) & 7u;
    float4 v0, v1;
    v0.x = acc[c][0] * kinv; v0.y = acc[c][1] * kinv; v0.z = acc[c][2] * kinv; v0.w = acc[c][3] * kinv;
    v1.x = acc[c][4] * kinv; v1.y = acc[c][5] * kinv; v1.z = acc[c][6] * kinv; v1.w = acc[c][7] * kinv;
    sts128(rowaddr + 16u * ((uint32_t)(2 * st) ^ sw), v0);
    sts128(rowaddr + 16u * ((uint32_t)(2 * st + 1) ^ sw), v1);
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0 && !(dbg & 2)) {
    if (role == 0) tma_store_4d_addr(wbase, &tm_gref, x0 + 32 * half, y0, c0 + cg * CG, b);
    else           tma_store_4d_addr(wbase, &tm_gfrm, x0 + 32 * half, y0, c0 + cg * CG, b);
    tma_store_commit();
    CVB_STAMP(20);
    tma_store_wait_read();   // shared memory must outlive the store's read
  }
  CVB_STAMP(21);
}
}  // namespace cvb


// =======================================================================================
// host dispatch
// =======================================================================================

int check_common(const float* const* frames, int F, int B, int C, int H, int W, int win) {
  if (!frames) return fail(B2F_EINVAL, "costvol: frames is NULL");
  if (F < 2 || F > kMaxFrames) return fail(B2F_EINVAL, "costvol: F=%d outside [2,%d]", F, kMaxFrames);
  if (B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "costvol: bad size B=%d C=%d H=%d W=%d", B, C, H, W);
  if (win <= 0 || (win & 1) == 0) return fail(B2F_EINVAL, "costvol: win=%d must be odd and positive", win);
  for (int f = 0; f < F; ++f) {
    if (!frames[f]) return fail(B2F_EINVAL, "costvol: frames[%d] is NULL", f);
    if (!aligned4(frames[f])) return fail(B2F_EALIGN, "costvol: frames[%d] misaligned", f);
  }
  return B2F_OK;
}

int grid_for(int64_t total, int threads) {
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <int A, int SGN, bool PIPE>
int launch_fwd_tma(const float* ref, const float* frm, float* out, int64_t obs, int B, int C, int H, int W,
                   float kdiv, int nsplit, cudaStream_t st) {
  using cfg = cvf::Cfg<A, PIPE>;
  const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)W, (uint64_t)W * H, (uint64_t)W * H * C};
  const uint32_t box_r[4] = {(uint32_t)cfg::RW, (uint32_t)cvf::TH, (uint32_t)cvf::CK, 1};
  const uint32_t box_f[4] = {(uint32_t)cfg::FW, (uint32_t)cfg::FR, (uint32_t)cvf::CK, 1};
  CUtensorMap tr, tf;
  int rc = make_tmap4(&tr, ref, dims, str, box_r);
  if (rc) return rc;
  if ((rc = make_tmap4(&tf, frm, dims, str, box_f))) return rc;
  CUtensorMap to = tr;   // PIPE only: view of the output for the staged TMA store
  bool warp_store = false;
  if (PIPE) {
    // per-warp stores: (x, y, window row iy, window column ix, batch) view, one 32 x 8 x 1 x 9 box per compute warp
    // (channel = ix*9 + iy); B2F_CVF_TILE_STORE=1 or a refused 5-D encode falls back to one 32 x 8 x 81 box per tile
    static const bool tile_store = [] { const char* e = getenv("B2F_CVF_TILE_STORE"); return e && e[0] == '1'; }();
    if (!tile_store) {
      const uint64_t hw = (uint64_t)W * H;
      const uint64_t odims5[5] = {(uint64_t)W, (uint64_t)H, 9, 9, (uint64_t)B};
      const uint64_t ostr5[4] = {(uint64_t)W, hw, 9 * hw, (uint64_t)obs};
      const uint32_t box5[5] = {(uint32_t)cfg::TW, (uint32_t)cvf::TH, 1, 9, 1};
      warp_store = make_tmap5(&to, out, odims5, ostr5, box5, true) == B2F_OK;
    }
    if (!warp_store) {
      const uint64_t odims[4] = {(uint64_t)W, (uint64_t)H, 81, (uint64_t)B};
      const uint64_t ostr[3] = {(uint64_t)W, (uint64_t)W * H, (uint64_t)obs};
      const uint32_t box_o[4] = {(uint32_t)cfg::TW, (uint32_t)cvf::TH, 81, 1};
      if ((rc = make_tmap4(&to, out, odims, ostr, box_o, true))) return rc;
    }
  }
  auto kern = cvf::costvol_fwd_tma<A, SGN, PIPE>;
  static thread_local int attr_dev = -1;
  int dev = 0;
  B2F_CUDA_TRY(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    B2F_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    attr_dev = dev;
  }
  const int nchunks = (C + cvf::CK - 1) / cvf::CK;
  const int cps = (nchunks + nsplit - 1) / nsplit;
  nsplit = (nchunks + cps - 1) / cps;   // no empty splits
  if (nsplit > 1)  // partial sums are accumulated: zero this instance's (B, 81, H, W) block first
    B2F_CUDA_TRY(cudaMemset2DAsync(out, (size_t)obs * 4, 0, (size_t)81 * H * W * 4, (size_t)B, st));
  const int ntx = (W + cfg::TW - 1) / cfg::TW, nty = (H + cvf::TH - 1) / cvf::TH;
  const int64_t ntiles = (int64_t)ntx * nty * B * nsplit;
  if (ntiles > 0x3fffffff) return fail(B2F_EINVAL, "costvol_forward: too many tiles");
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)num_sms() * cfg::CTAS_PER_SM);
  const int path = costvol_path();
  const int dbg = ((path >= 8 && path <= 10) ? path - 7 : 0) | (warp_store ? 4 : 0);   // 8: no arithmetic, 9: no stores, 10: neither; bit 2: per-warp stores
  kern<<<grid, cvf::THREADS, cfg::SMEM_BYTES, st>>>(tr, tf, to, out, obs, C, H, W, kdiv, nsplit, cps, ntx, nty,
                                                    (int)ntiles, dbg);
  B2F_CHECK_LAUNCH("costvol_fwd_tma");
  return B2F_OK;
}

template <int A, bool PIPE>
int launch_fwd_sgn(int sgn, const float* ref, const float* frm, float* out, int64_t obs, int B, int C, int H,
                   int W, float kdiv, int nsplit, cudaStream_t st) {
  return sgn > 0 ? launch_fwd_tma<A, 1, PIPE>(ref, frm, out, obs, B, C, H, W, kdiv, nsplit, st)
                 : launch_fwd_tma<A, -1, PIPE>(ref, frm, out, obs, B, C, H, W, kdiv, nsplit, st);
}

int64_t tiles_for(int A, int B, int H, int W) {
  return (int64_t)B * ((H + 7) / 8) * ((W + 4 * A - 1) / (4 * A));
}

}  // namespace
}  // namespace b2f

using namespace b2f;

extern "C" int b2f_costvol_forward(const float* const* frames, int F, int B, int C, int H, int W, int win,
                                   int fwd, float* out, int64_t out_batch_stride, b2f_stream_t stream) {
  int rc = check_common(frames, F, B, C, H, W, win);
  if (rc) return rc;
  if (!out) return fail(B2F_EINVAL, "costvol_forward: out is NULL");
  if (!aligned4(out)) return fail(B2F_EALIGN, "costvol_forward: out misaligned");
  const int64_t hw = (int64_t)H * W;
  const int64_t obs = out_batch_stride ? out_batch_stride : (int64_t)win * win * hw;
  if (obs < (int64_t)win * win * hw) return fail(B2F_EINVAL, "costvol_forward: out_batch_stride too small");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float kdiv = (float)(C * (F - 1));
  const int sgn = fwd ? 1 : -1;

  const int path = costvol_path();
  // 16 forces the tensor-core forward, 17 forbids it
  // (18 / 19: measurement modes of the tensor-core kernel without global stores / without TMEM loads: WRONG results)
  if ((path == 16 || path == 18 || path == 19) && F == 2 && win == 9 && (W % 4) == 0 && aligned16(frames[0]) && aligned16(frames[1]))
    return launch_costvol_fwd_tc(frames[0], frames[1], out, obs, B, C, H, W, kdiv, sgn, path == 18 ? 1 : (path == 19 ? 2 : 0), st);
  const bool tma_ok = path != 1 && F == 2 && win == 9 && (W % 4) == 0 && aligned16(frames[0]) &&
                      aligned16(frames[1]) && aligned16(out) && (obs % 4) == 0 && get_encode_fn() != nullptr;
  if (tma_ok) {
    // largest strip width whose grid still fills the machine
    // Measured on B200 (tools/time_cv.py): the 32-column tiles with two CTAs per SM beat the 64-column
    // ones at level 3 (60 vs 75 us); below one tile per SM the 16-column tiles win, and when even those do
    // not fill the machine the channel range is split across CTAs (path 5 forces a split in tests).
    int A = 0, nsplit = 1;
    const int sms = num_sms();
    if (path >= 2 && path <= 4) A = path == 4 ? 4 : 8;   // 2 and 3: 32-column tiles, 4: 16-column tiles
    else if (path >= 6 && path <= 10) A = 8;
    // 13 / 14 select the backward variant only: automatic forward
    else if (W >= 32 && tiles_for(8, B, H, W) >= sms) A = 8;
    else A = 4;
    if (A == 4 && (path == 0 || path == 5 || path >= 13)) {
      const int64_t t = tiles_for(4, B, H, W);
      const int nchunks = (C + cvf::CK - 1) / cvf::CK;
      if (2 * t <= sms || path == 5) nsplit = (int)std::min<int64_t>(nchunks, std::max<int64_t>(path == 5 ? 2 : 1, sms / t));
      if ((int64_t)B * nsplit > 65535) nsplit = 1;
    }
    // software-pipelined single-CTA form from 1.5 tiles per SM on (path 6 forces it, 7 forbids); measured with the
    // per-warp epilogue: level 4 of the benchmark (224 tiles on 148 SMs) 30.7-32.8 us against 32.8-34.8 us for the
    // two-CTAs-per-SM form, level 5 (64 tiles) 24.5 against 18.4 us
    const bool pipe = path == 6 || (path >= 8 && path <= 10) || (path != 7 && 2 * tiles_for(A, B, H, W) * nsplit >= 3 * (int64_t)sms);
    if (A == 8)
      return pipe ? launch_fwd_sgn<8, true>(sgn, frames[0], frames[1], out, obs, B, C, H, W, kdiv, nsplit, st)
                  : launch_fwd_sgn<8, false>(sgn, frames[0], frames[1], out, obs, B, C, H, W, kdiv, nsplit, st);
    return launch_fwd_sgn<4, false>(sgn, frames[0], frames[1], out, obs, B, C, H, W, kdiv, nsplit, st);
  }

  FramePtrs fp;
  for (int f = 0; f < kMaxFrames; ++f) fp.p[f] = f < F ? frames[f] : nullptr;
  const int64_t total = (int64_t)B * win * win * hw;
  costvol_fwd_generic<<<grid_for(total, 256), 256, 0, st>>>(fp, F, B, C, H, W, win, sgn, out, obs, kdiv);
  B2F_CHECK_LAUNCH("costvol_fwd_generic");
  return B2F_OK;
}

extern "C" int b2f_costvol_backward(const float* const* frames, int F, int B, int C, int H, int W, int win,
                                    int fwd, const float* gradOut, int64_t gradOut_batch_stride,
                                    float* const* gradFrames, b2f_stream_t stream) {
  int rc = check_common(frames, F, B, C, H, W, win);
  if (rc) return rc;
  if (!gradOut || !gradFrames) return fail(B2F_EINVAL, "costvol_backward: NULL gradOut/gradFrames");
  if (!aligned4(gradOut)) return fail(B2F_EALIGN, "costvol_backward: gradOut misaligned");
  const int64_t hw = (int64_t)H * W;
  const int64_t gbs = gradOut_batch_stride ? gradOut_batch_stride : (int64_t)win * win * hw;
  if (gbs < (int64_t)win * win * hw) return fail(B2F_EINVAL, "costvol_backward: gradOut_batch_stride too small");
  for (int f = 0; f < F; ++f)
    if (gradFrames[f] && !aligned4(gradFrames[f])) return fail(B2F_EALIGN, "costvol_backward: gradFrames[%d] misaligned", f);
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float kdiv = (float)(C * (F - 1));
  const int sgn = fwd ? 1 : -1;

  const int nroles = (gradFrames[0] ? 1 : 0) + ((F == 2 && gradFrames[1]) ? 1 : 0);
  const int path = costvol_path();
  bool tma_ok = path != 1 && F == 2 && win == 9 && (W % 4) == 0 && nroles > 0 &&
                aligned16(frames[0]) && aligned16(frames[1]) && aligned16(gradOut) && (gbs % 4) == 0 &&
                (!gradFrames[0] || aligned16(gradFrames[0])) && (!gradFrames[1] || aligned16(gradFrames[1])) &&
                get_encode_fn() != nullptr;
  if (tma_ok) {
    const int nchunk = (C + cvb::CC - 1) / cvb::CC;
    const int nty = (H + cvb::TH - 1) / cvb::TH;
    const int64_t per_col = (int64_t)B * nty * nchunk * nroles;
    const int64_t ctas32 = per_col * ((W + 31) / 32), ctas64 = per_col * ((W + 63) / 64);
    if ((ctas32 < 48 && path < 2) || ctas32 > 0x3fffffff || per_col / nty > 65535) tma_ok = false;
    if (tma_ok) {
      // variant: 13 / 14 / 15 force <32, 9> / <32, 2> / <64, 3>; automatic: small launches take the nine-slab
      // form, everything else <32, 2>.  The wide form is measured SLOWER (level 3: 119 vs 109 us, level 4: 70
      // vs 62 us) although it halves the TMA rows per pixel: with one CTA per SM the eight warps wait on the
      // same slab at the same time; it stays selectable for experiments (path 15) and is covered by the tests.
      int variant;   // 0: <32,2>, 1: <32,9>, 2: <64,3>
      const int64_t sms = num_sms();
      if (path == 13) variant = 1;
      else if (path == 14) variant = 0;
      else if (path == 15) variant = 2;
      else if (ctas32 <= sms * 3) variant = 1;
      else variant = 0;
      (void)ctas64;
      const int TWv = variant == 2 ? 64 : 32;
      const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
      const uint64_t str[3] = {(uint64_t)W, (uint64_t)hw, (uint64_t)hw * C};
      const uint32_t box_x[4] = {(uint32_t)(TWv + 12), (uint32_t)(cvb::XR / 2), (uint32_t)cvb::CG, 1};
      const uint64_t gdims[4] = {(uint64_t)W, (uint64_t)H, 81, (uint64_t)B};
      const uint64_t gstr[3] = {(uint64_t)W, (uint64_t)hw, (uint64_t)gbs};
      const uint32_t box_g0[4] = {(uint32_t)(TWv + 4), (uint32_t)cvb::TH, 1, 1};
      const uint32_t box_g1[4] = {(uint32_t)(TWv + 12), (uint32_t)cvb::TH, 1, 1};
      CUtensorMap tfrm, tref, tgo, tgo1;
      if ((rc = make_tmap4(&tfrm, frames[1], dims, str, box_x))) return rc;
      if ((rc = make_tmap4(&tref, frames[0], dims, str, box_x))) return rc;
      if ((rc = make_tmap4(&tgo, gradOut, gdims, gstr, box_g0))) return rc;
      if ((rc = make_tmap4(&tgo1, gradOut, gdims, gstr, box_g1))) return rc;
      // results: 32 x 8 x 8-channel boxes, 128-byte swizzled staging (a missing role reuses the other map)
      const uint32_t box_o[4] = {32, (uint32_t)cvb::TH, (uint32_t)cvb::CG, 1};
      CUtensorMap tgr, tgf;
      float* any = gradFrames[0] ? gradFrames[0] : gradFrames[1];
      if ((rc = make_tmap4(&tgr, gradFrames[0] ? gradFrames[0] : any, dims, str, box_o, true))) return rc;
      if ((rc = make_tmap4(&tgf, gradFrames[1] ? gradFrames[1] : any, dims, str, box_o, true))) return rc;
      static thread_local int attr_dev = -1;
      int dev = 0;
      B2F_CUDA_TRY(cudaGetDevice(&dev));
      if (attr_dev != dev) {
#define B2F_BWD_ATTR(TWX, NS)                                                                                      \
  B2F_CUDA_TRY(cudaFuncSetAttribute(cvb::costvol_bwd_tma<1, TWX, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    cvb::Cfg<TWX, NS>::SMEM_BYTES));                                               \
  B2F_CUDA_TRY(cudaFuncSetAttribute(cvb::costvol_bwd_tma<-1, TWX, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    cvb::Cfg<TWX, NS>::SMEM_BYTES))
        B2F_BWD_ATTR(32, 2);
        B2F_BWD_ATTR(32, 9);
        B2F_BWD_ATTR(64, 3);
#undef B2F_BWD_ATTR
        attr_dev = dev;
      }
      const int role0 = gradFrames[0] ? 0 : 1;
      const int bdbg = (path >= 8 && path <= 10) ? path - 7 : 0;   // 8: no arithmetic, 9: no stores, 10: neither
      dim3 grid((W + TWv - 1) / TWv, nty, B * nchunk * nroles);
#define B2F_BWD_LAUNCH(SG, TWX, NS)                                                                        \
  cvb::costvol_bwd_tma<SG, TWX, NS><<<grid, cvb::Cfg<TWX, NS>::THREADS, cvb::Cfg<TWX, NS>::SMEM_BYTES, st>>>( \
      tfrm, tref, tgo, tgo1, tgr, tgf, nroles, role0, nchunk, C, H, W, kdiv, bdbg)
      if (variant == 1) {
        if (sgn > 0) B2F_BWD_LAUNCH(1, 32, 9);
        else B2F_BWD_LAUNCH(-1, 32, 9);
      } else if (variant == 2) {
        if (sgn > 0) B2F_BWD_LAUNCH(1, 64, 3);
        else B2F_BWD_LAUNCH(-1, 64, 3);
      } else {
        if (sgn > 0) B2F_BWD_LAUNCH(1, 32, 2);
        else B2F_BWD_LAUNCH(-1, 32, 2);
      }
#undef B2F_BWD_LAUNCH
      B2F_CHECK_LAUNCH("costvol_bwd_tma");
      return B2F_OK;
    }
  }

  FramePtrs fp;
  for (int f = 0; f < kMaxFrames; ++f) fp.p[f] = f < F ? frames[f] : nullptr;
  const int64_t total = (int64_t)B * C * hw;
  for (int f = 0; f < F; ++f) {
    if (!gradFrames[f]) continue;
    costvol_bwd_generic<<<grid_for(total, 256), 256, 0, st>>>(fp, F, f, B, C, H, W, win, sgn, gradOut, gbs,
                                                              gradFrames[f], kdiv);
    B2F_CHECK_LAUNCH("costvol_bwd_generic");
  }
  return B2F_OK;
}

#ifdef B2F_CVB_TRACE
extern "C" __attribute__((visibility("default"))) int b2f_debug_cvb_trace(void* buf) {
  return (int)cudaMemcpyToSymbol(cvb::g_cvb_trace, &buf, sizeof(buf));
}
#endif
