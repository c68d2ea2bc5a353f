// nn.BilinearSamplerBHWD forward / backward for sm_100a.
//
// Semantics follow the reference's CUDA kernels (extras/stnbhwd/BilinearSamplerBHWD.cu), not its CPU
// code: pixel-unit offsets, grid channel 0 = x, 1 = y (Q2), clamp to the border, a tap at
// index W or H reads 0, no clamp derivative in the flow gradient (Q3).
//
//   xc = clamp(x + gx, 0, W-1); xi = floor(xc); wx = 1 - (xc - xi)        (getTopLeft, .cu:6-20)
//   out = wx wy TL + (1-wx) wy TR + wx (1-wy) BL + (1-wx)(1-wy) BR        (.cu:94-110)
//   gradImg[tap] += w_tap * gradOut  (in-bounds taps)                      (.cu:236-262)
//   gradGrid.x = -wy D_TL + wy D_TR - (1-wy) D_BL + (1-wy) D_BR            (.cu:287-295)
//   gradGrid.y = -wx D_TL + wx D_BL - (1-wx) D_TR + (1-wx) D_BR,  D_tap = sum_c img_tap * gradOut
//
// The geometry (add, clamp, floor, weight) is evaluated with exactly the reference's fp32 operations
// because floor() makes it discontinuous; the blend may contract to FMAs (well inside 1e-4).
//
// Launch shape: grid = (x tiles, output row, batch) so no thread does a 64-bit division.
//   * C % 4 == 0 (feature warps): 8 lanes share a pixel and stride over its float4 channel chunks --
//     one 128-byte line per tap per 32 channels; the four dot products of the flow gradient are
//     reduced with three shuffles; the image-gradient scatter is one red.global.add.v4.f32 per tap
//     and chunk (the reference issues one scalar atomicAdd per channel).
//   * C == 3 (image warps): one thread per pixel.  Forward stages the 128 x 3 outputs of a block in
//     shared memory so every store instruction writes 512 contiguous bytes.  Backward: the TL/TR
//     (and BL/BR) taps are 6 contiguous floats, scattered with 2-3 vector reductions chosen by the
//     address alignment instead of 6 scalar ones -- the kernel is bound by the SM's RED issue rate.
//   * any other C: one thread per pixel, scalar.
#include "tma.cuh"

namespace b2f {
namespace {

struct Geo {
  int xi, yi;
  float wx, wy;
  bool rin, bin;  // right / bottom neighbour inside the image
};

__device__ __forceinline__ void top_left(float off, int idx, int size, int& point, float& weight) {
  float xc = __fadd_rn(off, (float)idx);
  if (xc < 0.f) xc = 0.f;
  if (xc > (float)(size - 1)) xc = (float)(size - 1);
  const float fl = floorf(xc);
  point = (int)fl;
  weight = __fsub_rn(1.f, __fsub_rn(xc, fl));
}

__device__ __forceinline__ Geo geometry(float gx, float gy, int xo, int yo, int H, int W) {
  Geo g;
  top_left(gx, xo, W, g.xi, g.wx);
  top_left(gy, yo, H, g.yi, g.wy);
  g.rin = g.xi + 1 <= W - 1;
  g.bin = g.yi + 1 <= H - 1;
  return g;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// Scatter-add 6 contiguous floats (two adjacent 3-channel pixels) with the fewest reductions the
// address alignment allows.  `pad_ok`: p[6] is inside the buffer (a zero may be added to it).
__device__ __forceinline__ void red_add6(float* p, const float (&v)[6], bool pad_ok) {
  const unsigned a = (unsigned)((reinterpret_cast<uintptr_t>(p) >> 2) & 3u);
  if (a == 0) {
    red_add_v4(p, v[0], v[1], v[2], v[3]);
    red_add_v2(p + 4, v[4], v[5]);
  } else if (a == 2) {
    red_add_v2(p, v[0], v[1]);
    red_add_v4(p + 2, v[2], v[3], v[4], v[5]);
  } else if (a == 3) {
    red_add(p, v[0]);
    red_add_v4(p + 1, v[1], v[2], v[3], v[4]);
    red_add(p + 5, v[5]);
  } else {  // a == 1
    red_add(p, v[0]);
    red_add_v2(p + 1, v[1], v[2]);
    if (pad_ok) {
      red_add_v4(p + 3, v[3], v[4], v[5], 0.f);
    } else {
      red_add_v2(p + 3, v[3], v[4]);
      red_add(p + 5, v[5]);
    }
  }
}
__device__ __forceinline__ void red_add3(float* p, float a, float b, float c) {
  if ((reinterpret_cast<uintptr_t>(p) & 7u) == 0) {
    red_add_v2(p, a, b);
    red_add(p + 2, c);
  } else {
    red_add(p, a);
    red_add_v2(p + 1, b, c);
  }
}

// Six contiguous floats (two adjacent 3-channel pixels) starting at a 4-byte aligned address, fetched as
// 2-3 aligned 16-byte loads plus a register selection instead of 6 scalar loads: the C = 3 kernels are
// bound by L1/LSU transactions (a scalar warp request touches ~28 sectors), not by bytes.  [lo, hi) is the
// tensor's extent; chunks that would cross it fall back to scalar loads.
// The loads and the selection are separate calls so that a thread can have both tap rows (and the taps of
// several pixels) in flight before it consumes any of them -- these kernels are latency-bound.
struct Raw6 {
  float4 c0, c1, c2;
  unsigned a;
};
__device__ __forceinline__ void load6_issue(Raw6& r, const float* p, const float* lo, const float* hi) {
  r.a = (unsigned)((reinterpret_cast<uintptr_t>(p) >> 2) & 3u);
  const float* base = p - r.a;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  if (base >= lo && base + 12 <= hi) {
    r.c0 = __ldg(reinterpret_cast<const float4*>(base));
    r.c1 = __ldg(reinterpret_cast<const float4*>(base) + 1);
    r.c2 = (r.a == 3) ? __ldg(reinterpret_cast<const float4*>(base) + 2) : zero;
  } else {  // first / last bytes of the tensor: element-wise, never outside [lo, hi)
    float t[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) t[i] = (base + i >= lo && base + i < hi) ? __ldg(base + i) : 0.f;
    r.c0 = make_float4(t[0], t[1], t[2], t[3]);
    r.c1 = make_float4(t[4], t[5], t[6], t[7]);
    r.c2 = make_float4(t[8], t[9], t[10], t[11]);
  }
}
__device__ __forceinline__ void load6_select(const Raw6& r, float (&v)[6]) {
  const float4 c0 = r.c0, c1 = r.c1, c2 = r.c2;
  if (r.a == 0) { v[0] = c0.x; v[1] = c0.y; v[2] = c0.z; v[3] = c0.w; v[4] = c1.x; v[5] = c1.y; }
  else if (r.a == 1) { v[0] = c0.y; v[1] = c0.z; v[2] = c0.w; v[3] = c1.x; v[4] = c1.y; v[5] = c1.z; }
  else if (r.a == 2) { v[0] = c0.z; v[1] = c0.w; v[2] = c1.x; v[3] = c1.y; v[4] = c1.z; v[5] = c1.w; }
  else { v[0] = c0.w; v[1] = c1.x; v[2] = c1.y; v[3] = c1.z; v[4] = c1.w; v[5] = c2.x; }
}

constexpr int LPP = 8;           // lanes per pixel on the vector path
constexpr int VEC_THREADS = 256; // 32 pixels per block
constexpr int PIX_THREADS = 128; // one thread per pixel

// ---- forward, C % 4 == 0 ---------------------------------------------------------------
// Each 8-lane group handles two pixels of the row (x and x + 32): both grid values are fetched first,
// then all eight tap loads are in flight together -- the kernel is latency-bound, not byte-bound.
constexpr int VEC_PPT = 2;
__global__ void __launch_bounds__(VEC_THREADS, 4)
warp_fwd_vec4(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
              int H, int W, int C, int Hg, int Wg) {
  const int xa = blockIdx.x * (VEC_THREADS / LPP) * VEC_PPT + (threadIdx.x >> 3);
  const int yo = blockIdx.y, b = blockIdx.z;
  const int sub = threadIdx.x & (LPP - 1);
  const int nch4 = C >> 2;
  const size_t row = ((size_t)b * Hg + yo) * Wg;
  int xo[VEC_PPT];
  bool live[VEC_PPT];
  float2 gxy[VEC_PPT];
#pragma unroll
  for (int p = 0; p < VEC_PPT; ++p) {
    xo[p] = xa + p * (VEC_THREADS / LPP);
    live[p] = xo[p] < Wg;
    gxy[p] = live[p] ? __ldg(reinterpret_cast<const float2*>(grid) + row + xo[p]) : make_float2(0.f, 0.f);
  }
  Geo g[VEC_PPT];
  const float4* tl[VEC_PPT];
#pragma unroll
  for (int p = 0; p < VEC_PPT; ++p) {
    g[p] = geometry(gxy[p].x, gxy[p].y, live[p] ? xo[p] : 0, yo, H, W);
    tl[p] = reinterpret_cast<const float4*>(img + (((size_t)b * H + g[p].yi) * W + g[p].xi) * C);
  }
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t rowq = (size_t)W * nch4;
  for (int q = sub; q < nch4; q += LPP) {
    float4 a[VEC_PPT], c[VEC_PPT], d[VEC_PPT], e[VEC_PPT];
#pragma unroll
    for (int p = 0; p < VEC_PPT; ++p) {
      const bool rin = g[p].rin && live[p], bin = g[p].bin && live[p];
      a[p] = live[p] ? __ldg(tl[p] + q) : zero;
      c[p] = rin ? __ldg(tl[p] + nch4 + q) : zero;
      d[p] = bin ? __ldg(tl[p] + rowq + q) : zero;
      e[p] = (rin && bin) ? __ldg(tl[p] + rowq + nch4 + q) : zero;
    }
#pragma unroll
    for (int p = 0; p < VEC_PPT; ++p) {
      if (!live[p]) continue;
      const float w_tl = g[p].wx * g[p].wy, w_tr = (1.f - g[p].wx) * g[p].wy;
      const float w_bl = g[p].wx * (1.f - g[p].wy), w_br = (1.f - g[p].wx) * (1.f - g[p].wy);
      float4 v;
      v.x = w_tl * a[p].x + w_tr * c[p].x + w_bl * d[p].x + w_br * e[p].x;
      v.y = w_tl * a[p].y + w_tr * c[p].y + w_bl * d[p].y + w_br * e[p].y;
      v.z = w_tl * a[p].z + w_tr * c[p].z + w_bl * d[p].z + w_br * e[p].z;
      v.w = w_tl * a[p].w + w_tr * c[p].w + w_bl * d[p].w + w_br * e[p].w;
      __stcs(reinterpret_cast<float4*>(out + (row + xo[p]) * C) + q, v);
    }
  }
}

// ---- forward, C == 3: one thread per pixel, stores staged through shared memory ------------
__global__ void __launch_bounds__(PIX_THREADS)
warp_fwd_c3(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
            int H, int W, int Hg, int Wg, const float* img_end) {
  __shared__ float s_out[PIX_THREADS * 3];
  const int x0 = blockIdx.x * PIX_THREADS;
  const int xo = x0 + threadIdx.x;
  const int yo = blockIdx.y, b = blockIdx.z;
  const size_t row = ((size_t)b * Hg + yo) * Wg;
  if (xo < Wg) {
    const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + row + xo);
    const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
    const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
    const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
    const float* tl = img + (((size_t)b * H + g.yi) * W + g.xi) * 3;
    // a right tap at index W (a bottom tap at index H) contributes 0: give it a zero weight and let the
    // six-float load run over into the neighbouring row, which is inside the tensor or guarded by load6
    Raw6 rt, rb;
    load6_issue(rt, tl, img, img_end);
    load6_issue(rb, g.bin ? tl + (size_t)W * 3 : tl, img, img_end);
    float top[6], bot[6];
    load6_select(rt, top);
    load6_select(rb, bot);
    const bool both = g.rin && g.bin;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      s_out[threadIdx.x * 3 + c] = w_tl * top[c] + w_tr * (g.rin ? top[3 + c] : 0.f) +
                                   w_bl * (g.bin ? bot[c] : 0.f) + w_br * (both ? bot[3 + c] : 0.f);
  }
  __syncthreads();
  const int n = min(PIX_THREADS, Wg - x0) * 3;
  float* o = out + (row + x0) * 3;
  for (int i = threadIdx.x; i < n; i += PIX_THREADS) __stcs(o + i, s_out[i]);
}

// ---- forward, any C: one thread per pixel ---------------------------------------------------
__global__ void __launch_bounds__(PIX_THREADS)
warp_fwd_scalar(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
                int H, int W, int C, int Hg, int Wg) {
  const int xo = blockIdx.x * PIX_THREADS + threadIdx.x;
  if (xo >= Wg) return;
  const int yo = blockIdx.y, b = blockIdx.z;
  const size_t pix = ((size_t)b * Hg + yo) * Wg + xo;
  const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + pix);
  const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
  const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
  const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
  const float* tl = img + (((size_t)b * H + g.yi) * W + g.xi) * C;
  const float* bl = tl + (size_t)W * C;
  const bool both = g.rin && g.bin;
  for (int c = 0; c < C; ++c) {
    const float a = __ldg(tl + c);
    const float t = g.rin ? __ldg(tl + C + c) : 0.f;
    const float d = g.bin ? __ldg(bl + c) : 0.f;
    const float e = both ? __ldg(bl + C + c) : 0.f;
    out[pix * C + c] = w_tl * a + w_tr * t + w_bl * d + w_br * e;
  }
}

// ---- backward, C % 4 == 0 -------------------------------------------------------------
// Like the forward, an 8-lane group handles two pixels (x and x + 32), and per channel chunk ALL ten loads
// (gradOut + four taps, both pixels) are issued before the first reduction: a red.global is a compiler
// memory barrier, so interleaving "load tap, reduce, load next tap" serialises four L2 round trips per
// pixel (measured: 76 % of the stall samples on the long scoreboard, 25 % issue utilisation).
template <bool ONLY_GRID>
__global__ void __launch_bounds__(VEC_THREADS)
warp_bwd_vec4(const float* __restrict__ img, const float* __restrict__ grid, const float* __restrict__ gout,
              float* __restrict__ gimg, float* __restrict__ ggrid, int H, int W, int C, int Hg, int Wg) {
  // whole warps stay alive for the shuffles: out-of-range pixels are just not "live"
  const int xa = blockIdx.x * (VEC_THREADS / LPP) * VEC_PPT + (threadIdx.x >> 3);
  const int b = blockIdx.z;
  const int sub = threadIdx.x & (LPP - 1);
  const int nch4 = C >> 2;
  int xo[VEC_PPT];
  bool live[VEC_PPT];
#pragma unroll
  for (int p = 0; p < VEC_PPT; ++p) {
    xo[p] = xa + p * (VEC_THREADS / LPP);
    live[p] = xo[p] < Wg;
  }
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t rowq = (size_t)W * nch4;
  const size_t rowC = (size_t)W * C;
  // rows blockIdx.y, blockIdx.y + gridDim.y, ... with the next row's grid values prefetched (see the forward)
  float2 gnext[VEC_PPT];
#pragma unroll
  for (int p = 0; p < VEC_PPT; ++p)
    gnext[p] = live[p] ? __ldg(reinterpret_cast<const float2*>(grid) + ((size_t)b * Hg + blockIdx.y) * Wg + xo[p])
                       : make_float2(0.f, 0.f);
  for (int yo = blockIdx.y; yo < Hg; yo += gridDim.y) {
    const size_t row = ((size_t)b * Hg + yo) * Wg;
    float2 gxy[VEC_PPT];
#pragma unroll
    for (int p = 0; p < VEC_PPT; ++p) {
      gxy[p] = gnext[p];
      if (yo + (int)gridDim.y < Hg && live[p])
        gnext[p] = __ldg(reinterpret_cast<const float2*>(grid) + row + (size_t)gridDim.y * Wg + xo[p]);
    }
    Geo g[VEC_PPT];
    const float4* tl[VEC_PPT];
    const float4* go[VEC_PPT];
    float* gi[VEC_PPT];
    float dot[VEC_PPT][4];
#pragma unroll
    for (int p = 0; p < VEC_PPT; ++p) {
      g[p] = geometry(gxy[p].x, gxy[p].y, live[p] ? xo[p] : 0, yo, H, W);
      const size_t a0 = (((size_t)b * H + g[p].yi) * W + g[p].xi) * C;
      tl[p] = reinterpret_cast<const float4*>(img + a0);
      go[p] = reinterpret_cast<const float4*>(gout + (row + (live[p] ? xo[p] : 0)) * C);
      gi[p] = ONLY_GRID ? nullptr : gimg + a0;
#pragma unroll
      for (int t = 0; t < 4; ++t) dot[p][t] = 0.f;
    }
    for (int q = sub; q < nch4; q += LPP) {
      float4 v[VEC_PPT], a[VEC_PPT], c[VEC_PPT], d[VEC_PPT], e[VEC_PPT];
#pragma unroll
      for (int p = 0; p < VEC_PPT; ++p) {
        const bool rin = g[p].rin && live[p], bin = g[p].bin && live[p];
        v[p] = live[p] ? __ldcs(go[p] + q) : zero;
        a[p] = live[p] ? __ldg(tl[p] + q) : zero;
        c[p] = rin ? __ldg(tl[p] + nch4 + q) : zero;
        d[p] = bin ? __ldg(tl[p] + rowq + q) : zero;
        e[p] = (rin && bin) ? __ldg(tl[p] + rowq + nch4 + q) : zero;
      }
#pragma unroll
      for (int p = 0; p < VEC_PPT; ++p) {
        if (!live[p]) continue;
        const float4 w = v[p];
        dot[p][0] += a[p].x * w.x + a[p].y * w.y + a[p].z * w.z + a[p].w * w.w;
        dot[p][1] += c[p].x * w.x + c[p].y * w.y + c[p].z * w.z + c[p].w * w.w;
        dot[p][2] += d[p].x * w.x + d[p].y * w.y + d[p].z * w.z + d[p].w * w.w;
        dot[p][3] += e[p].x * w.x + e[p].y * w.y + e[p].z * w.z + e[p].w * w.w;
        if (!ONLY_GRID) {
          const float w_tl = g[p].wx * g[p].wy, w_tr = (1.f - g[p].wx) * g[p].wy;
          const float w_bl = g[p].wx * (1.f - g[p].wy), w_br = (1.f - g[p].wx) * (1.f - g[p].wy);
          float* o = gi[p] + 4 * q;
          red_add_v4(o, w_tl * w.x, w_tl * w.y, w_tl * w.z, w_tl * w.w);
          if (g[p].rin) red_add_v4(o + C, w_tr * w.x, w_tr * w.y, w_tr * w.z, w_tr * w.w);
          if (g[p].bin) red_add_v4(o + rowC, w_bl * w.x, w_bl * w.y, w_bl * w.z, w_bl * w.w);
          if (g[p].rin && g[p].bin) red_add_v4(o + rowC + C, w_br * w.x, w_br * w.y, w_br * w.z, w_br * w.w);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < VEC_PPT; ++p) {
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1)
#pragma unroll
        for (int t = 0; t < 4; ++t) dot[p][t] += __shfl_xor_sync(0xffffffffu, dot[p][t], o);
      if (live[p] && sub == 0) {
        // taps outside the image contributed zero (their loads were replaced by zeros)
        float2 r;
        r.x = -g[p].wy * dot[p][0] + g[p].wy * dot[p][1] - (1.f - g[p].wy) * dot[p][2] + (1.f - g[p].wy) * dot[p][3];
        r.y = -g[p].wx * dot[p][0] + g[p].wx * dot[p][2] - (1.f - g[p].wx) * dot[p][1] + (1.f - g[p].wx) * dot[p][3];
        reinterpret_cast<float2*>(ggrid)[row + xo[p]] = r;
      }
    }
  }
}

// ---- backward, C == 3: one thread per pixel, alignment-aware vector reductions ---------------
template <bool ONLY_GRID>
__global__ void __launch_bounds__(PIX_THREADS)
warp_bwd_c3(const float* __restrict__ img, const float* __restrict__ grid, const float* __restrict__ gout,
            float* __restrict__ gimg, float* __restrict__ ggrid, int H, int W, int Hg, int Wg,
            size_t gimg_elems) {
  __shared__ float s_go[PIX_THREADS * 3];
  const int x0 = blockIdx.x * PIX_THREADS;
  const int xo = x0 + threadIdx.x;
  const int yo = blockIdx.y, b = blockIdx.z;
  const size_t row = ((size_t)b * Hg + yo) * Wg;
  // coalesced read of the block's 128 x 3 gradOut floats
  const int n = min(PIX_THREADS, Wg - x0) * 3;
  const float* gsrc = gout + (row + x0) * 3;
  for (int i = threadIdx.x; i < n; i += PIX_THREADS) s_go[i] = __ldcs(gsrc + i);
  __syncthreads();
  if (xo >= Wg) return;
  const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + row + xo);
  const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
  const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
  const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
  const size_t a0 = (((size_t)b * H + g.yi) * W + g.xi) * 3;
  const size_t rowC = (size_t)W * 3;
  const float v0 = s_go[threadIdx.x * 3], v1 = s_go[threadIdx.x * 3 + 1], v2 = s_go[threadIdx.x * 3 + 2];
  const float* img_end = img + gimg_elems;
  Raw6 rt, rb;
  load6_issue(rt, img + a0, img, img_end);
  load6_issue(rb, g.bin ? img + a0 + rowC : img + a0, img, img_end);
  float top[6], bot[6];
  load6_select(rt, top);
  load6_select(rb, bot);
  const float d_tl = top[0] * v0 + top[1] * v1 + top[2] * v2;
  const float d_tr = g.rin ? top[3] * v0 + top[4] * v1 + top[5] * v2 : 0.f;
  const float d_bl = g.bin ? bot[0] * v0 + bot[1] * v1 + bot[2] * v2 : 0.f;
  const float d_br = (g.rin && g.bin) ? bot[3] * v0 + bot[4] * v1 + bot[5] * v2 : 0.f;
  if (!ONLY_GRID) {
    if (g.rin) {
      const float top[6] = {w_tl * v0, w_tl * v1, w_tl * v2, w_tr * v0, w_tr * v1, w_tr * v2};
      red_add6(gimg + a0, top, a0 + 6 < gimg_elems);
      if (g.bin) {
        const float bot[6] = {w_bl * v0, w_bl * v1, w_bl * v2, w_br * v0, w_br * v1, w_br * v2};
        red_add6(gimg + a0 + rowC, bot, a0 + rowC + 6 < gimg_elems);
      }
    } else {
      red_add3(gimg + a0, w_tl * v0, w_tl * v1, w_tl * v2);
      if (g.bin) red_add3(gimg + a0 + rowC, w_bl * v0, w_bl * v1, w_bl * v2);
    }
  }
  float2 r;
  r.x = -g.wy * d_tl + g.wy * d_tr - (1.f - g.wy) * d_bl + (1.f - g.wy) * d_br;
  r.y = -g.wx * d_tl + g.wx * d_bl - (1.f - g.wx) * d_tr + (1.f - g.wx) * d_br;
  reinterpret_cast<float2*>(ggrid)[row + xo] = r;
}

// ---- C == 3, lean forms (the model's image warps) ----------------------------------------------------------
// The kernels above spend ~320 issue slots per pixel (4-way alignment branches, 64-bit index arithmetic, extent
// checks, an unrolled copy loop): ncu shows them issue-bound at 0.6 IPC per scheduler with the memory system a
// quarter busy, and a smooth flow field runs no faster than an i.i.d. one.  These forms do the same arithmetic
// with 32-bit offsets and no divergent code: the six contiguous floats of a tap row are fetched as two (three
// when the run starts at float 3 of a 16-byte chunk) aligned 128-bit loads and moved into place by a two-stage
// predicated barrel shift; the scatter is the mirror image -- the six products are shifted into 16-byte chunks
// padded with zeros and leave as two red.global.add.v4.f32 (plus one scalar when the run spills into a third
// chunk): 4.5 reduction operations per pixel instead of 5-6, and adding +0 to a neighbour is exact.
// Preconditions (checked by the dispatcher): 16-byte aligned tensors, B*H*W*3 a multiple of 4 (every chunk that
// holds a valid float is then inside the tensor) and below 2^31, Wg a multiple of 4.
struct Chunks9 {
  float4 c0, c1;
  float c2;
};
__device__ __forceinline__ Chunks9 load_chunks(const float* __restrict__ img, unsigned a, unsigned total) {
  const unsigned base = a & ~3u;
  const float4* p = reinterpret_cast<const float4*>(img + base);
  Chunks9 r;
  r.c0 = __ldg(p);
  r.c1 = (base + 8u <= total) ? __ldg(p + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
  r.c2 = ((a & 3u) == 3u && base + 12u <= total) ? __ldg(img + base + 8) : 0.f;
  return r;
}
// v[i] = chunk floats [o + i], o = a & 3
__device__ __forceinline__ void shift_out(const Chunks9& r, unsigned o, float (&v)[6]) {
  const float c[9] = {r.c0.x, r.c0.y, r.c0.z, r.c0.w, r.c1.x, r.c1.y, r.c1.z, r.c1.w, r.c2};
  const bool p2 = (o & 2u) != 0, p1 = (o & 1u) != 0;
  float t[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) t[i] = p2 ? c[i + 2] : c[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = p1 ? t[i + 1] : t[i];
}
// scatter-add v[0..6) at float offset a: the inverse shift, then aligned vector reductions
__device__ __forceinline__ void red_chunks(float* __restrict__ gimg, unsigned a, unsigned total, const float (&v)[6]) {
  const unsigned o = a & 3u, base = a & ~3u;
  const bool p2 = (o & 2u) != 0, p1 = (o & 1u) != 0;
  float u[7];
  u[0] = p1 ? 0.f : v[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) u[i] = p1 ? v[i - 1] : v[i];
  u[6] = p1 ? v[5] : 0.f;
  float s[9];
  s[0] = p2 ? 0.f : u[0];
  s[1] = p2 ? 0.f : u[1];
#pragma unroll
  for (int i = 2; i < 7; ++i) s[i] = p2 ? u[i - 2] : u[i];
  s[7] = p2 ? u[5] : 0.f;
  s[8] = p2 ? u[6] : 0.f;
  float* p = gimg + base;
  red_add_v4(p, s[0], s[1], s[2], s[3]);
  if (base + 8u <= total) red_add_v4(p + 4, s[4], s[5], s[6], s[7]);
  if (o == 3u && base + 12u <= total) red_add(p + 8, s[8]);
}

__global__ void __launch_bounds__(PIX_THREADS)
warp_fwd_c3_lean(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
                 int H, int W, int Hg, int Wg, unsigned total) {
  __shared__ __align__(16) float s_out[PIX_THREADS * 3];
  const int x0 = blockIdx.x * PIX_THREADS;
  const int xo = x0 + threadIdx.x;
  const int yo = blockIdx.y, b = blockIdx.z;
  const unsigned row = ((unsigned)b * Hg + yo) * Wg;
  if (xo < Wg) {
    const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + row + xo);
    const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
    const unsigned a = (((unsigned)b * H + g.yi) * W + g.xi) * 3u;
    const unsigned ab = g.bin ? a + (unsigned)W * 3u : a;
    const Chunks9 rt = load_chunks(img, a, total);
    const Chunks9 rb = load_chunks(img, ab, total);
    const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
    const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
    float top[6], bot[6];
    shift_out(rt, a & 3u, top);
    shift_out(rb, ab & 3u, bot);
    const bool both = g.rin && g.bin;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      s_out[threadIdx.x * 3 + c] = w_tl * top[c] + w_tr * (g.rin ? top[3 + c] : 0.f) +
                                   w_bl * (g.bin ? bot[c] : 0.f) + w_br * (both ? bot[3 + c] : 0.f);
  }
  __syncthreads();
  // 128 x 3 floats leave as 96 x 16 bytes (Wg % 4 == 0: the row segment starts and ends on a 16-byte boundary)
  const int nq = (min(PIX_THREADS, Wg - x0) * 3) >> 2;
  if ((int)threadIdx.x < nq)
    __stcs(reinterpret_cast<float4*>(out + (size_t)(row + x0) * 3) + threadIdx.x,
           reinterpret_cast<const float4*>(s_out)[threadIdx.x]);
}

// ---- C == 3, windowed forward (the large image warps) ---------------------------------------------------------
// With scattered flow (the benchmark's i.i.d. N(0, 4 px)) every lane of a warp gathers from a different image row:
// a global gather costs one L1 wavefront per lane and request, and warp_fwd_c3_lean is bound by L1 wavefronts
// (~4.5 per pixel: 54 us at 8 x 448 x 1024, 33 % of the HBM roofline) with DRAM a third busy.  Here a CTA owns a
// 32 x 64 output tile and has the TMA stage the source window of the tile +- 12 pixels (58 rows x 96 pixels: one box of the
// image viewed as 16-pixel groups, 66.8 KB) in shared memory while the threads fetch their flow values; the twelve tap values of a
// pixel are then shared-memory loads (bank conflicts, ~3.5-way on random addresses, instead of 32-way wavefronts).  The
// window position is STATIC, so the TMA is in flight from the first instruction; a pixel whose taps leave the window
// (|flow| > 12 px vertically, 15 px horizontally: 0.3 % of an N(0, 4) field) takes the global path of the lean kernel -- the result never depends on
// the window.  Three CTAs per SM: one tile's fill runs under the others' gathers.  Rows / columns of the window
// outside the image read as zero or as a neighbouring batch item's rows; neither is ever used (the coordinates are
// clamped to the image, and a tap at index W or H is masked by rin / bin).  Same arithmetic expressions as
// warp_fwd_c3_lean: bit-identical output.
namespace win {
constexpr int TH = 32, TW = 64, THREADS = 256, PPT = TH * TW / THREADS;
constexpr int RX = 16, RY = 12;             // window margins: 16 columns (the box starts on a 16-pixel group), 12 rows
constexpr int WR = TH + 2 * RY + 2;         // 58 window rows
constexpr int WPX = TW + 2 * RX;            // 96 window columns = 6 groups of 16 pixels (48 floats)
constexpr int PITCH = WPX * 3;              // 288 floats: ONE box {48 floats, 6 groups, 58 rows} of the image viewed as
                                            // (48, W / 16, B H) lands the window with contiguous rows
constexpr int BOX_ELEMS = WR * PITCH;
constexpr int SMEM_BYTES = BOX_ELEMS * 4 + (THREADS / 32) * 96 * 4 + 64 + 128;
static_assert((BOX_ELEMS * 4) % 128 == 0, "TMA destination alignment");
}  // namespace win

__global__ void __launch_bounds__(win::THREADS, 3)
warp_fwd_c3_win(const __grid_constant__ CUtensorMap tm_img, const float* __restrict__ img,
                const float* __restrict__ grid, float* __restrict__ out, int H, int W, unsigned total) {
  using namespace win;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const uint32_t box0 = smem_u32(smem);
  float* s_out = reinterpret_cast<float*>(smem) + BOX_ELEMS;                 // [warp][96]
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_out + (THREADS / 32) * 96);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
  const int wx0 = x0 - RX, wy0 = y0 - RY;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(bar, BOX_ELEMS * 4);
    tma_load_4d(smem, &tm_img, 0, wx0 / 16, b * H + wy0, 0, bar);
  }
  // thread -> pixels: warp w walks rows w, w + 8, ...; lanes cover 32 consecutive columns of one half-row
  Geo g[PPT];
  bool live[PPT];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int idx = k * THREADS + tid;                 // 0 .. 2047
    const int ry = idx >> 6, rx = idx & 63;
    const int yo = y0 + ry, xo = x0 + rx;
    live[k] = yo < H && xo < W;
    float2 gxy = make_float2(0.f, 0.f);
    if (live[k]) gxy = __ldg(reinterpret_cast<const float2*>(grid) + ((unsigned)b * H + yo) * W + xo);
    g[k] = geometry(gxy.x, gxy.y, live[k] ? xo : 0, live[k] ? yo : 0, H, W);
  }
  __syncthreads();                                     // the barrier is initialised
  mbar_wait(bar, 0);
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int idx = k * THREADS + tid;
    const int ry = idx >> 6, rx = idx & 63;
    const Geo& q = g[k];
    float top[6], bot[6];
    const int cx = q.xi - wx0, cy = q.yi - wy0;
    const bool inwin = cx >= 0 && cx + 1 < WPX && cy >= 0 && cy + 1 < WR;
    if (live[k]) {
      if (inwin) {
        const uint32_t a = box0 + 4u * (uint32_t)(cy * PITCH + cx * 3);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(top[i]) : "r"(a + 4u * i));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bot[i]) : "r"(a + 4u * (PITCH + i)));
        }
      } else {
        const unsigned a = (((unsigned)b * H + q.yi) * W + q.xi) * 3u;
        const unsigned ab = q.bin ? a + (unsigned)W * 3u : a;
        const Chunks9 rt = load_chunks(img, a, total);
        const Chunks9 rb = load_chunks(img, ab, total);
        shift_out(rt, a & 3u, top);
        shift_out(rb, ab & 3u, bot);
      }
    }
    const float w_tl = q.wx * q.wy, w_tr = (1.f - q.wx) * q.wy;
    const float w_bl = q.wx * (1.f - q.wy), w_br = (1.f - q.wx) * (1.f - q.wy);
    const bool both = q.rin && q.bin;
    float* so = s_out + warp * 96 + lane * 3;
    __syncwarp();
    if (live[k]) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        so[c] = w_tl * top[c] + w_tr * (q.rin ? top[3 + c] : 0.f) + w_bl * (q.bin ? bot[c] : 0.f) +
                w_br * (both ? bot[3 + c] : 0.f);
    }
    __syncwarp();
    // the warp's 32 pixels are 96 contiguous floats of one output row (W % 4 == 0: 16-byte aligned segments)
    const int yo = y0 + ry, xw = x0 + (rx & 32);
    if (yo < H && lane < 24) {
      const int nq = (min(32, W - xw) * 3) >> 2;       // float4 count of the live part of the segment
      if (lane < nq)
        __stcs(reinterpret_cast<float4*>(out + (size_t)(((unsigned)b * H + yo) * W + xw) * 3) + lane,
               reinterpret_cast<const float4*>(s_out + warp * 96)[lane]);
    }
  }
}

template <bool ONLY_GRID>
__global__ void __launch_bounds__(PIX_THREADS)
warp_bwd_c3_lean(const float* __restrict__ img, const float* __restrict__ grid, const float* __restrict__ gout,
                 float* __restrict__ gimg, float* __restrict__ ggrid, int H, int W, int Hg, int Wg, unsigned total) {
  const int xo = blockIdx.x * PIX_THREADS + threadIdx.x;
  if (xo >= Wg) return;
  const int yo = blockIdx.y, b = blockIdx.z;
  const unsigned pix = ((unsigned)b * Hg + yo) * Wg + xo;
  const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + pix);
  const float* gp = gout + (size_t)pix * 3;
  const float v0 = __ldcs(gp), v1 = __ldcs(gp + 1), v2 = __ldcs(gp + 2);
  const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
  const unsigned a = (((unsigned)b * H + g.yi) * W + g.xi) * 3u;
  const unsigned ab = g.bin ? a + (unsigned)W * 3u : a;
  const Chunks9 rt = load_chunks(img, a, total);
  const Chunks9 rb = load_chunks(img, ab, total);
  const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
  const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
  float top[6], bot[6];
  shift_out(rt, a & 3u, top);
  shift_out(rb, ab & 3u, bot);
  const float d_tl = top[0] * v0 + top[1] * v1 + top[2] * v2;
  const float d_tr = g.rin ? top[3] * v0 + top[4] * v1 + top[5] * v2 : 0.f;
  const float d_bl = g.bin ? bot[0] * v0 + bot[1] * v1 + bot[2] * v2 : 0.f;
  const float d_br = (g.rin && g.bin) ? bot[3] * v0 + bot[4] * v1 + bot[5] * v2 : 0.f;
  float2 r;
  r.x = -g.wy * d_tl + g.wy * d_tr - (1.f - g.wy) * d_bl + (1.f - g.wy) * d_br;
  r.y = -g.wx * d_tl + g.wx * d_bl - (1.f - g.wx) * d_tr + (1.f - g.wx) * d_br;
  reinterpret_cast<float2*>(ggrid)[pix] = r;
  if (!ONLY_GRID) {
    // a tap outside the image receives nothing (BilinearSamplerBHWD.cu:236-262): its products are replaced by
    // zeros, which the padded reductions add harmlessly to the floats that happen to follow
    const float tr0 = g.rin ? w_tr * v0 : 0.f, tr1 = g.rin ? w_tr * v1 : 0.f, tr2 = g.rin ? w_tr * v2 : 0.f;
    const float vt[6] = {w_tl * v0, w_tl * v1, w_tl * v2, tr0, tr1, tr2};
    red_chunks(gimg, a, total, vt);
    if (g.bin) {
      const float br0 = g.rin ? w_br * v0 : 0.f, br1 = g.rin ? w_br * v1 : 0.f, br2 = g.rin ? w_br * v2 : 0.f;
      const float vb[6] = {w_bl * v0, w_bl * v1, w_bl * v2, br0, br1, br2};
      red_chunks(gimg, ab, total, vb);
    }
  }
}

// ---- C == 3, windowed backward: the four taps of the flow gradient come from the staged window (as in the forward);
// the image-gradient scatter stays a global reduction (shared-memory fp32 atomics are CAS loops on sm_100a) ----------
template <bool ONLY_GRID>
__global__ void __launch_bounds__(win::THREADS, 3)
warp_bwd_c3_win(const __grid_constant__ CUtensorMap tm_img, const float* __restrict__ img, const float* __restrict__ grid,
                const float* __restrict__ gout, float* __restrict__ gimg, float* __restrict__ ggrid, int H, int W,
                unsigned total) {
  using namespace win;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const uint32_t box0 = smem_u32(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem) + BOX_ELEMS + (THREADS / 32) * 96);
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
  const int wx0 = x0 - RX, wy0 = y0 - RY;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(bar, BOX_ELEMS * 4);
    tma_load_4d(smem, &tm_img, 0, wx0 / 16, b * H + wy0, 0, bar);
  }
  Geo g[PPT];
  bool live[PPT];
  float go[PPT][3];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int idx = k * THREADS + tid;
    const int yo = y0 + (idx >> 6), xo = x0 + (idx & 63);
    live[k] = yo < H && xo < W;
    float2 gxy = make_float2(0.f, 0.f);
    go[k][0] = go[k][1] = go[k][2] = 0.f;
    if (live[k]) {
      const unsigned pix = ((unsigned)b * H + yo) * W + xo;
      gxy = __ldg(reinterpret_cast<const float2*>(grid) + pix);
      const float* gp = gout + (size_t)pix * 3;
      go[k][0] = __ldcs(gp); go[k][1] = __ldcs(gp + 1); go[k][2] = __ldcs(gp + 2);
    }
    g[k] = geometry(gxy.x, gxy.y, live[k] ? xo : 0, live[k] ? yo : 0, H, W);
  }
  __syncthreads();
  mbar_wait(bar, 0);
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    if (!live[k]) continue;
    const int idx = k * THREADS + tid;
    const int yo = y0 + (idx >> 6), xo = x0 + (idx & 63);
    const unsigned pix = ((unsigned)b * H + yo) * W + xo;
    const Geo& q = g[k];
    const float v0 = go[k][0], v1 = go[k][1], v2 = go[k][2];
    float top[6], bot[6];
    const int cx = q.xi - wx0, cy = q.yi - wy0;
    const unsigned a = (((unsigned)b * H + q.yi) * W + q.xi) * 3u;
    const unsigned ab = q.bin ? a + (unsigned)W * 3u : a;
    if (cx >= 0 && cx + 1 < WPX && cy >= 0 && cy + 1 < WR) {
      const uint32_t sa = box0 + 4u * (uint32_t)(cy * PITCH + cx * 3);
      const uint32_t sb = sa + (q.bin ? 4u * PITCH : 0u);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(top[i]) : "r"(sa + 4u * i));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bot[i]) : "r"(sb + 4u * i));
      }
    } else {
      const Chunks9 rt = load_chunks(img, a, total);
      const Chunks9 rb = load_chunks(img, ab, total);
      shift_out(rt, a & 3u, top);
      shift_out(rb, ab & 3u, bot);
    }
    const float w_tl = q.wx * q.wy, w_tr = (1.f - q.wx) * q.wy;
    const float w_bl = q.wx * (1.f - q.wy), w_br = (1.f - q.wx) * (1.f - q.wy);
    const float d_tl = top[0] * v0 + top[1] * v1 + top[2] * v2;
    const float d_tr = q.rin ? top[3] * v0 + top[4] * v1 + top[5] * v2 : 0.f;
    const float d_bl = q.bin ? bot[0] * v0 + bot[1] * v1 + bot[2] * v2 : 0.f;
    const float d_br = (q.rin && q.bin) ? bot[3] * v0 + bot[4] * v1 + bot[5] * v2 : 0.f;
    float2 r;
    r.x = -q.wy * d_tl + q.wy * d_tr - (1.f - q.wy) * d_bl + (1.f - q.wy) * d_br;
    r.y = -q.wx * d_tl + q.wx * d_bl - (1.f - q.wx) * d_tr + (1.f - q.wx) * d_br;
    reinterpret_cast<float2*>(ggrid)[pix] = r;
    if (!ONLY_GRID) {
      const float tr0 = q.rin ? w_tr * v0 : 0.f, tr1 = q.rin ? w_tr * v1 : 0.f, tr2 = q.rin ? w_tr * v2 : 0.f;
      const float vt[6] = {w_tl * v0, w_tl * v1, w_tl * v2, tr0, tr1, tr2};
      red_chunks(gimg, a, total, vt);
      if (q.bin) {
        const float br0 = q.rin ? w_br * v0 : 0.f, br1 = q.rin ? w_br * v1 : 0.f, br2 = q.rin ? w_br * v2 : 0.f;
        const float vb[6] = {w_bl * v0, w_bl * v1, w_bl * v2, br0, br1, br2};
        red_chunks(gimg, ab, total, vb);
      }
    }
  }
}

// ---- backward, any C: one thread per pixel --------------------------------------------
template <bool ONLY_GRID>
__global__ void __launch_bounds__(PIX_THREADS)
warp_bwd_scalar(const float* __restrict__ img, const float* __restrict__ grid, const float* __restrict__ gout,
                float* __restrict__ gimg, float* __restrict__ ggrid, int H, int W, int C, int Hg, int Wg) {
  const int xo = blockIdx.x * PIX_THREADS + threadIdx.x;
  if (xo >= Wg) return;
  const int yo = blockIdx.y, b = blockIdx.z;
  const size_t pix = ((size_t)b * Hg + yo) * Wg + xo;
  const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + pix);
  const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
  const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
  const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
  const size_t a0 = (((size_t)b * H + g.yi) * W + g.xi) * C;
  const size_t rowC = (size_t)W * C;
  const bool both = g.rin && g.bin;
  float d_tl = 0.f, d_tr = 0.f, d_bl = 0.f, d_br = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = __ldg(gout + pix * C + c);
    d_tl += __ldg(img + a0 + c) * v;
    if (!ONLY_GRID) red_add(gimg + a0 + c, w_tl * v);
    if (g.rin) {
      d_tr += __ldg(img + a0 + C + c) * v;
      if (!ONLY_GRID) red_add(gimg + a0 + C + c, w_tr * v);
    }
    if (g.bin) {
      d_bl += __ldg(img + a0 + rowC + c) * v;
      if (!ONLY_GRID) red_add(gimg + a0 + rowC + c, w_bl * v);
    }
    if (both) {
      d_br += __ldg(img + a0 + rowC + C + c) * v;
      if (!ONLY_GRID) red_add(gimg + a0 + rowC + C + c, w_br * v);
    }
  }
  float2 r;
  r.x = -g.wy * d_tl + g.wy * d_tr - (1.f - g.wy) * d_bl + (1.f - g.wy) * d_br;
  r.y = -g.wx * d_tl + g.wx * d_bl - (1.f - g.wx) * d_tr + (1.f - g.wx) * d_br;
  reinterpret_cast<float2*>(ggrid)[pix] = r;
}

// ---- warpingUnit fused: BDHW in, BDHW out -------------------------------------------
// models/pwc.lua:68-73 wraps the sampler in four nn.Transpose copies (image and flow to BHWD, result back) behind a
// nn.MulConstant on the flow (:402-408, :441-446): three extra passes over every warped tensor.  These kernels
// read the network's planar tensors directly: out[b,c,y,x] = bilinear(img[b,c], x + s*flow[b,0,y,x],
// y + s*flow[b,1,y,x]).  One thread per pixel; the geometry is evaluated once per pixel (the BHWD kernels evaluate it
// in every channel lane) and the channel loop gathers four taps per plane -- neighbouring lanes read neighbouring
// addresses when the flow is smooth, which is what the decoder's bilinearly up-sampled flow is.  s*flow is rounded
// to fp32 before the pixel coordinate is added, as MulConstant followed by getTopLeft does.
// CQ > 1 (small pyramid levels, too few pixels to fill the machine with one thread per pixel): the block is
// (128 / CQ pixels) x (CQ channel ranges); C % (CQ * UNROLL) == 0.
template <int UNROLL, int CQ>
__global__ void __launch_bounds__(PIX_THREADS)
warp_bdhw_fwd(const float* __restrict__ img, const float* __restrict__ flow, float scale, float* __restrict__ out,
              int C, int H, int W) {
  constexpr int PX = PIX_THREADS / CQ;
  const int x = blockIdx.x * PX + (threadIdx.x % PX);
  const int cq = threadIdx.x / PX;
  if (x >= W) return;
  const int y = blockIdx.y;
  const unsigned b = blockIdx.z, hw = (unsigned)H * W, pix = (unsigned)y * W + x;
  const float* fl = flow + (size_t)b * 2 * hw + pix;
  const Geo g = geometry(__fmul_rn(__ldg(fl), scale), __fmul_rn(__ldg(fl + hw), scale), x, y, H, W);
  const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
  const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
  const unsigned o_tl = (unsigned)g.yi * W + g.xi;
  const unsigned d_r = g.rin ? 1u : 0u, d_b = g.bin ? (unsigned)W : 0u;   // absent taps re-read TL and are zeroed
  const bool both = g.rin && g.bin;
  const float* ip = img + (size_t)b * C * hw + o_tl;
  float* op = out + (size_t)b * C * hw + pix;
  int c = CQ > 1 ? cq * (C / CQ) : 0;
  const int C_end = CQ > 1 ? c + C / CQ : C;
  for (; c + UNROLL <= C_end; c += UNROLL) {
    float tl[UNROLL], tr[UNROLL], bl[UNROLL], br[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const float* p = ip + (size_t)(c + u) * hw;
      tl[u] = __ldg(p);
      tr[u] = __ldg(p + d_r);
      bl[u] = __ldg(p + d_b);
      br[u] = __ldg(p + d_b + d_r);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      __stcs(op + (size_t)(c + u) * hw, w_tl * tl[u] + w_tr * (g.rin ? tr[u] : 0.f) + w_bl * (g.bin ? bl[u] : 0.f) +
                                            w_br * (both ? br[u] : 0.f));
  }
  for (; c < C_end; ++c) {
    const float* p = ip + (size_t)c * hw;
    const float tl = __ldg(p), tr = __ldg(p + d_r), bl = __ldg(p + d_b), br = __ldg(p + d_b + d_r);
    __stcs(op + (size_t)c * hw, w_tl * tl + w_tr * (g.rin ? tr : 0.f) + w_bl * (g.bin ? bl : 0.f) + w_br * (both ? br : 0.f));
  }
}

// gradImg (planar, pre-zeroed by the caller, NULL = flow gradient only) is scattered with scalar reductions --
// with a smooth flow the 32 lanes of a warp hit one or two lines per tap; the four dot products of the flow
// gradient accumulate over the channel loop in the pixel's own thread (no cross-lane reduction), and the result is
// multiplied by s (MulConstant's backward) and stored planar.
template <int UNROLL, bool ONLY_GRID, int CQ>
__global__ void __launch_bounds__(PIX_THREADS)
warp_bdhw_bwd(const float* __restrict__ img, const float* __restrict__ flow, float scale,
              const float* __restrict__ gout, float* __restrict__ gimg, float* __restrict__ gflow, int C, int H, int W) {
  constexpr int PX = PIX_THREADS / CQ;
  __shared__ float s_dot[CQ > 1 ? CQ : 1][4][PX];
  const int lx = threadIdx.x % PX, cq = threadIdx.x / PX;
  const int x = blockIdx.x * PX + lx;
  const bool live = x < W;
  if (CQ == 1 && !live) return;
  if (CQ > 1 && !live) {   // keeps the barrier below uniform
    __syncthreads();
    return;
  }
  const int y = blockIdx.y;
  const unsigned b = blockIdx.z, hw = (unsigned)H * W, pix = (unsigned)y * W + x;
  const float* fl = flow + (size_t)b * 2 * hw + pix;
  const Geo g = geometry(__fmul_rn(__ldg(fl), scale), __fmul_rn(__ldg(fl + hw), scale), x, y, H, W);
  const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
  const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
  const unsigned o_tl = (unsigned)g.yi * W + g.xi;
  const unsigned d_r = g.rin ? 1u : 0u, d_b = g.bin ? (unsigned)W : 0u;
  const bool both = g.rin && g.bin;
  const float* ip = img + (size_t)b * C * hw + o_tl;
  const float* gp = gout + (size_t)b * C * hw + pix;
  float* gi = ONLY_GRID ? nullptr : gimg + (size_t)b * C * hw + o_tl;
  float d_tl = 0.f, d_tr = 0.f, d_bl = 0.f, d_br = 0.f;
  int c = CQ > 1 ? cq * (C / CQ) : 0;
  const int C_end = CQ > 1 ? c + C / CQ : C;
  for (; c + UNROLL <= C_end; c += UNROLL) {
    float v[UNROLL], tl[UNROLL], tr[UNROLL], bl[UNROLL], br[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const float* p = ip + (size_t)(c + u) * hw;
      v[u] = __ldcs(gp + (size_t)(c + u) * hw);
      tl[u] = __ldg(p);
      tr[u] = __ldg(p + d_r);
      bl[u] = __ldg(p + d_b);
      br[u] = __ldg(p + d_b + d_r);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      d_tl += tl[u] * v[u];
      d_tr += tr[u] * v[u];
      d_bl += bl[u] * v[u];
      d_br += br[u] * v[u];
      if (!ONLY_GRID) {
        float* q = gi + (size_t)(c + u) * hw;
        red_add(q, w_tl * v[u]);
        if (g.rin) red_add(q + 1, w_tr * v[u]);
        if (g.bin) red_add(q + W, w_bl * v[u]);
        if (both) red_add(q + W + 1, w_br * v[u]);
      }
    }
  }
  for (; c < C_end; ++c) {
    const float* p = ip + (size_t)c * hw;
    const float v = __ldcs(gp + (size_t)c * hw);
    d_tl += __ldg(p) * v;
    d_tr += __ldg(p + d_r) * v;
    d_bl += __ldg(p + d_b) * v;
    d_br += __ldg(p + d_b + d_r) * v;
    if (!ONLY_GRID) {
      float* q = gi + (size_t)c * hw;
      red_add(q, w_tl * v);
      if (g.rin) red_add(q + 1, w_tr * v);
      if (g.bin) red_add(q + W, w_bl * v);
      if (both) red_add(q + W + 1, w_br * v);
    }
  }
  if (CQ > 1) {   // fixed-order sum of the channel ranges' partial dot products
    s_dot[cq][0][lx] = d_tl; s_dot[cq][1][lx] = d_tr; s_dot[cq][2][lx] = d_bl; s_dot[cq][3][lx] = d_br;
    __syncthreads();
    if (cq != 0) return;
    d_tl = d_tr = d_bl = d_br = 0.f;
#pragma unroll
    for (int q = 0; q < CQ; ++q) {
      d_tl += s_dot[q][0][lx]; d_tr += s_dot[q][1][lx]; d_bl += s_dot[q][2][lx]; d_br += s_dot[q][3][lx];
    }
  }
  // a tap outside the image contributes nothing to the dot products (BilinearSamplerBHWD.cu:236-262)
  if (!g.rin) d_tr = 0.f;
  if (!g.bin) d_bl = 0.f;
  if (!both) d_br = 0.f;
  const float gx = -g.wy * d_tl + g.wy * d_tr - (1.f - g.wy) * d_bl + (1.f - g.wy) * d_br;
  const float gy = -g.wx * d_tl + g.wx * d_bl - (1.f - g.wx) * d_tr + (1.f - g.wx) * d_br;
  float* gf = gflow + (size_t)b * 2 * hw + pix;
  gf[0] = gx * scale;
  gf[hw] = gy * scale;
}

// gridDim.y of the vector backward kernel: a block walks every gridDim.y-th row (the forward measured
// slower with the row loop: its extra registers cost a resident block).  Enough blocks for ~2-3 waves of
// resident CTAs (2-3 per SM), at most 8 rows per block so the tail stays short.
int vec_rows_grid(int Wg, int Hg, int B) {
  const int64_t cols = (int64_t)((Wg + 32 * VEC_PPT - 1) / (32 * VEC_PPT)) * B;
  const int64_t want = (int64_t)num_sms() * 6;
  int64_t gy = (want + cols - 1) / cols;
  const int64_t gmin = (Hg + 7) / 8;
  if (gy < gmin) gy = gmin;
  if (gy > Hg) gy = Hg;
  return (int)gy;
}

// small levels of the fused warpingUnit: four channel ranges per pixel when one thread per pixel cannot fill the SMs
bool bdhw_split(int B, int C, int H, int W) { return C % 16 == 0 && (int64_t)B * H * W < (int64_t)num_sms() * 1024; }

// preconditions of the lean C = 3 kernels (see there); B2F_WARP_C3_LEGACY=1 forces the older kernels (experiments)
bool c3_lean_ok(int B, int H, int W, int Hg, int Wg) {
  static const bool legacy = [] { const char* e = getenv("B2F_WARP_C3_LEGACY"); return e && e[0] == '1'; }();
  const size_t total = (size_t)B * H * W * 3, gpix = (size_t)B * Hg * Wg;
  return !legacy && (total & 3u) == 0 && total < (1ull << 31) && gpix * 3 < (1ull << 31) && (Wg & 3) == 0;
}

// the windowed forward: the lean preconditions, output grid = image grid, rows of 16-byte multiples (TMA), at least a
// few waves of 32 x 64 tiles; B2F_WARP_C3_WIN=0 switches it off (experiments)
bool c3_win_ok(int B, int H, int W, int Hg, int Wg) {
  static const bool off = [] { const char* e = getenv("B2F_WARP_C3_WIN"); return e && e[0] == '0'; }();
  if (off || !c3_lean_ok(B, H, W, Hg, Wg) || H != Hg || W != Wg || (W & 15) != 0) return false;
  if ((int64_t)B * H > (1ll << 30)) return false;
  const int64_t tiles = (int64_t)B * ((H + win::TH - 1) / win::TH) * ((W + win::TW - 1) / win::TW);
  return tiles >= 2 * (int64_t)num_sms() && get_encode_fn() != nullptr;
}

// the (B, H, W, 3) image as (48 floats = 16 pixels, W / 16 groups, B H rows): one box is the whole window
int win_tmap(CUtensorMap* tm, const float* img, int B, int H, int W) {
  const uint64_t dims[4] = {48, (uint64_t)W / 16, (uint64_t)B * H, 1};
  const uint64_t str[3] = {48, (uint64_t)W * 3, (uint64_t)W * 3 * B * H};
  const uint32_t box[4] = {48, win::WPX / 16, (uint32_t)win::WR, 1};
  return make_tmap4(tm, img, dims, str, box);
}

int check_args(const float* img, const float* grid, int B, int H, int W, int C, int Hg, int Wg) {
  if (!img || !grid) return fail(B2F_EINVAL, "warp: NULL img/grid");
  if (B < 0 || H <= 0 || W <= 0 || C <= 0 || Hg <= 0 || Wg <= 0)
    return fail(B2F_EINVAL, "warp: bad size B=%d H=%d W=%d C=%d Hg=%d Wg=%d", B, H, W, C, Hg, Wg);
  if (B > 65535 || Hg > 65535) return fail(B2F_EINVAL, "warp: B and Hg must be <= 65535 (grid y/z limits)");
  if (!aligned4(img)) return fail(B2F_EALIGN, "warp: img misaligned");
  if ((reinterpret_cast<uintptr_t>(grid) & 7u) != 0) return fail(B2F_EALIGN, "warp: grid must be 8-byte aligned");
  return B2F_OK;
}

}  // namespace
}  // namespace b2f

using namespace b2f;

extern "C" int b2f_warp_bhwd_forward(const float* img, const float* grid, float* out, int B, int H, int W,
                                     int C, int Hg, int Wg, b2f_stream_t stream) {
  int rc = check_args(img, grid, B, H, W, C, Hg, Wg);
  if (rc) return rc;
  if (!out) return fail(B2F_EINVAL, "warp_forward: out is NULL");
  if (!aligned4(out)) return fail(B2F_EALIGN, "warp_forward: out misaligned");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((C & 3) == 0 && aligned16(img) && aligned16(out)) {
    dim3 grid_dim((Wg + 32 * VEC_PPT - 1) / (32 * VEC_PPT), Hg, B);
    warp_fwd_vec4<<<grid_dim, VEC_THREADS, 0, st>>>(img, grid, out, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_fwd_vec4");
  } else if (C == 3) {
    dim3 grid_dim((Wg + PIX_THREADS - 1) / PIX_THREADS, Hg, B);
    if (c3_win_ok(B, H, W, Hg, Wg) && aligned16(img) && aligned16(out)) {
      CUtensorMap tm;
      if ((rc = win_tmap(&tm, img, B, H, W))) return rc;
      static thread_local int attr_dev = -1;
      int dev = 0;
      B2F_CUDA_TRY(cudaGetDevice(&dev));
      if (attr_dev != dev) {
        B2F_CUDA_TRY(cudaFuncSetAttribute(warp_fwd_c3_win, cudaFuncAttributeMaxDynamicSharedMemorySize, win::SMEM_BYTES));
        attr_dev = dev;
      }
      dim3 gw((W + win::TW - 1) / win::TW, (H + win::TH - 1) / win::TH, B);
      warp_fwd_c3_win<<<gw, win::THREADS, win::SMEM_BYTES, st>>>(tm, img, grid, out, H, W, (unsigned)((size_t)B * H * W * 3));
      B2F_CHECK_LAUNCH("warp_fwd_c3_win");
    } else if (c3_lean_ok(B, H, W, Hg, Wg) && aligned16(img) && aligned16(out)) {
      warp_fwd_c3_lean<<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, out, H, W, Hg, Wg, (unsigned)((size_t)B * H * W * 3));
      B2F_CHECK_LAUNCH("warp_fwd_c3_lean");
    } else {
      warp_fwd_c3<<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, out, H, W, Hg, Wg, img + (size_t)B * H * W * 3);
      B2F_CHECK_LAUNCH("warp_fwd_c3");
    }
  } else {
    dim3 grid_dim((Wg + PIX_THREADS - 1) / PIX_THREADS, Hg, B);
    warp_fwd_scalar<<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, out, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_fwd_scalar");
  }
  return B2F_OK;
}

extern "C" int b2f_warp_bhwd_backward(const float* img, const float* grid, const float* gradOut,
                                      float* gradImg, float* gradGrid, int B, int H, int W, int C, int Hg,
                                      int Wg, b2f_stream_t stream) {
  int rc = check_args(img, grid, B, H, W, C, Hg, Wg);
  if (rc) return rc;
  if (!gradOut || !gradGrid) return fail(B2F_EINVAL, "warp_backward: NULL gradOut/gradGrid");
  if (!aligned4(gradOut) || (gradImg && !aligned4(gradImg))) return fail(B2F_EALIGN, "warp_backward: misaligned");
  if ((reinterpret_cast<uintptr_t>(gradGrid) & 7u) != 0) return fail(B2F_EALIGN, "warp_backward: gradGrid must be 8-byte aligned");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool only = gradImg == nullptr;
  if ((C & 3) == 0 && aligned16(img) && aligned16(gradOut) && (only || aligned16(gradImg))) {
    dim3 grid_dim((Wg + 32 * VEC_PPT - 1) / (32 * VEC_PPT), vec_rows_grid(Wg, Hg, B), B);
    if (only) warp_bwd_vec4<true><<<grid_dim, VEC_THREADS, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, H, W, C, Hg, Wg);
    else warp_bwd_vec4<false><<<grid_dim, VEC_THREADS, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_bwd_vec4");
  } else if (C == 3) {
    dim3 grid_dim((Wg + PIX_THREADS - 1) / PIX_THREADS, Hg, B);
    const size_t elems = (size_t)B * H * W * 3;
    // the windowed form pays only for the flow-gradient-only backward (45 vs 53 us at 8 x 448 x 1024, i.i.d. 4 px flow):
    // with the image-gradient scatter in the same kernel it measured 111 vs 113 us on scattered flow and 94-98 vs 84-88 us
    // on smooth flow -- the reductions, not the gathers, bound that kernel, and eight pixels per thread serialise them
    if (only && c3_win_ok(B, H, W, Hg, Wg) && aligned16(img)) {
      CUtensorMap tm;
      if ((rc = win_tmap(&tm, img, B, H, W))) return rc;
      static thread_local int attr_dev = -1;
      int dev = 0;
      B2F_CUDA_TRY(cudaGetDevice(&dev));
      if (attr_dev != dev) {
        B2F_CUDA_TRY(cudaFuncSetAttribute(warp_bwd_c3_win<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, win::SMEM_BYTES));
        B2F_CUDA_TRY(cudaFuncSetAttribute(warp_bwd_c3_win<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, win::SMEM_BYTES));
        attr_dev = dev;
      }
      dim3 gw((W + win::TW - 1) / win::TW, (H + win::TH - 1) / win::TH, B);
      if (only) warp_bwd_c3_win<true><<<gw, win::THREADS, win::SMEM_BYTES, st>>>(tm, img, grid, gradOut, gradImg, gradGrid, H, W, (unsigned)elems);
      else warp_bwd_c3_win<false><<<gw, win::THREADS, win::SMEM_BYTES, st>>>(tm, img, grid, gradOut, gradImg, gradGrid, H, W, (unsigned)elems);
      B2F_CHECK_LAUNCH("warp_bwd_c3_win");
    } else if (c3_lean_ok(B, H, W, Hg, Wg) && aligned16(img) && (only || aligned16(gradImg))) {
      if (only) warp_bwd_c3_lean<true><<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, H, W, Hg, Wg, (unsigned)elems);
      else warp_bwd_c3_lean<false><<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, H, W, Hg, Wg, (unsigned)elems);
      B2F_CHECK_LAUNCH("warp_bwd_c3_lean");
    } else {
      if (only) warp_bwd_c3<true><<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, H, W, Hg, Wg, elems);
      else warp_bwd_c3<false><<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, H, W, Hg, Wg, elems);
      B2F_CHECK_LAUNCH("warp_bwd_c3");
    }
  } else {
    dim3 grid_dim((Wg + PIX_THREADS - 1) / PIX_THREADS, Hg, B);
    if (only) warp_bwd_scalar<true><<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, H, W, C, Hg, Wg);
    else warp_bwd_scalar<false><<<grid_dim, PIX_THREADS, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_bwd_scalar");
  }
  return B2F_OK;
}

// warpingUnit(I, F) of models/pwc.lua:68-73 with the nn.MulConstant that feeds it (:402-408, :441-446) folded in.
extern "C" int b2f_warp_bdhw_forward(const float* img, const float* flow, float flow_scale, float* out, int B,
                                     int C, int H, int W, b2f_stream_t stream) {
  if (!img || !flow || !out) return fail(B2F_EINVAL, "warp_bdhw_forward: NULL pointer");
  if (B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "warp_bdhw: bad size B=%d C=%d H=%d W=%d", B, C, H, W);
  if (B > 65535 || H > 65535) return fail(B2F_EINVAL, "warp_bdhw: B and H must be <= 65535 (grid y/z limits)");
  if ((size_t)H * W >= (1ull << 31)) return fail(B2F_EINVAL, "warp_bdhw: H*W must be below 2^31");
  if (!aligned4(img) || !aligned4(flow) || !aligned4(out)) return fail(B2F_EALIGN, "warp_bdhw_forward: misaligned");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bdhw_split(B, C, H, W)) {
    dim3 grid_dim((W + 31) / 32, H, B);
    warp_bdhw_fwd<4, 4><<<grid_dim, PIX_THREADS, 0, st>>>(img, flow, flow_scale, out, C, H, W);
  } else {
    dim3 grid_dim((W + PIX_THREADS - 1) / PIX_THREADS, H, B);
    if (C % 4 == 0) warp_bdhw_fwd<4, 1><<<grid_dim, PIX_THREADS, 0, st>>>(img, flow, flow_scale, out, C, H, W);
    else warp_bdhw_fwd<3, 1><<<grid_dim, PIX_THREADS, 0, st>>>(img, flow, flow_scale, out, C, H, W);
  }
  B2F_CHECK_LAUNCH("warp_bdhw_fwd");
  return B2F_OK;
}

extern "C" int b2f_warp_bdhw_backward(const float* img, const float* flow, float flow_scale, const float* gradOut,
                                      float* gradImg, float* gradFlow, int B, int C, int H, int W,
                                      b2f_stream_t stream) {
  if (!img || !flow || !gradOut || !gradFlow) return fail(B2F_EINVAL, "warp_bdhw_backward: NULL pointer");
  if (B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "warp_bdhw: bad size B=%d C=%d H=%d W=%d", B, C, H, W);
  if (B > 65535 || H > 65535) return fail(B2F_EINVAL, "warp_bdhw: B and H must be <= 65535 (grid y/z limits)");
  if ((size_t)H * W >= (1ull << 31)) return fail(B2F_EINVAL, "warp_bdhw: H*W must be below 2^31");
  if (!aligned4(img) || !aligned4(flow) || !aligned4(gradOut) || !aligned4(gradFlow) || (gradImg && !aligned4(gradImg)))
    return fail(B2F_EALIGN, "warp_bdhw_backward: misaligned");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool only = gradImg == nullptr;
  const bool split = bdhw_split(B, C, H, W);
  dim3 grid_dim(split ? (W + 31) / 32 : (W + PIX_THREADS - 1) / PIX_THREADS, H, B);
#define B2F_BDHW_BWD(U, Q) do { if (only) warp_bdhw_bwd<U, true, Q><<<grid_dim, PIX_THREADS, 0, st>>>(img, flow, flow_scale, gradOut, gradImg, gradFlow, C, H, W); \
                                else warp_bdhw_bwd<U, false, Q><<<grid_dim, PIX_THREADS, 0, st>>>(img, flow, flow_scale, gradOut, gradImg, gradFlow, C, H, W); } while (0)
  if (split) B2F_BDHW_BWD(4, 4);
  else if (C % 4 == 0) B2F_BDHW_BWD(4, 1);
  else B2F_BDHW_BWD(3, 1);
#undef B2F_BDHW_BWD
  B2F_CHECK_LAUNCH("warp_bdhw_bwd");
  return B2F_OK;
}
