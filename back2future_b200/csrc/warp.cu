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
// Mapping: BHWD keeps a pixel's C channels contiguous, so for C % 4 == 0 eight lanes share a pixel and
// stride over its float4 channel chunks (one 128-byte line per tap per 32 channels); the four dot
// products are reduced with three shuffles.  Other C (the C = 3 image warps) use one thread per
// (pixel, channel) forward -- consecutive threads write consecutive floats -- and one thread per pixel
// backward.  The image-gradient scatter uses fire-and-forget reductions (red.global.add, .v4 for the
// vector path), like the reference's atomicAdd but 4 channels per instruction.
#include "common.cuh"

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
__device__ __forceinline__ void red_add(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

constexpr int LPP = 8;  // lanes per pixel on the vector path

// ---- forward, C % 4 == 0 ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
warp_fwd_vec4(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
              int B, int H, int W, int C, int Hg, int Wg) {
  const int64_t npix = (int64_t)B * Hg * Wg;
  const int sub = threadIdx.x & (LPP - 1);
  const int nch4 = C >> 2;
  for (int64_t pix = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPP; pix < npix;
       pix += (int64_t)gridDim.x * blockDim.x / LPP) {
    const int xo = (int)(pix % Wg);
    const int yo = (int)((pix / Wg) % Hg);
    const int b = (int)(pix / ((int64_t)Wg * Hg));
    const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + pix);
    const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
    const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
    const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
    const float4* tl = reinterpret_cast<const float4*>(img + (((int64_t)b * H + g.yi) * W + g.xi) * C);
    const float4* tr = tl + nch4;
    const float4* bl = tl + (int64_t)W * nch4;
    const float4* br = bl + nch4;
    float4* o = reinterpret_cast<float4*>(out + pix * C);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = sub; q < nch4; q += LPP) {
      const float4 a = __ldg(tl + q);
      const float4 c = g.rin ? __ldg(tr + q) : zero;
      const float4 d = g.bin ? __ldg(bl + q) : zero;
      const float4 e = (g.rin && g.bin) ? __ldg(br + q) : zero;
      float4 v;
      v.x = w_tl * a.x + w_tr * c.x + w_bl * d.x + w_br * e.x;
      v.y = w_tl * a.y + w_tr * c.y + w_bl * d.y + w_br * e.y;
      v.z = w_tl * a.z + w_tr * c.z + w_bl * d.z + w_br * e.z;
      v.w = w_tl * a.w + w_tr * c.w + w_bl * d.w + w_br * e.w;
      o[q] = v;
    }
  }
}

// ---- forward, any C: one thread per output float -----------------------------------------
template <int CT>  // CT > 0: compile-time channel count, 0: runtime
__global__ void __launch_bounds__(256)
warp_fwd_scalar(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
                int B, int H, int W, int Crt, int Hg, int Wg) {
  const int C = CT > 0 ? CT : Crt;
  const int64_t total = (int64_t)B * Hg * Wg * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = idx / C;
    const int c = (int)(idx - pix * C);
    const int xo = (int)(pix % Wg);
    const int yo = (int)((pix / Wg) % Hg);
    const int b = (int)(pix / ((int64_t)Wg * Hg));
    const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + pix);
    const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
    const float* tl = img + (((int64_t)b * H + g.yi) * W + g.xi) * C + c;
    const float a = __ldg(tl);
    const float t = g.rin ? __ldg(tl + C) : 0.f;
    const float d = g.bin ? __ldg(tl + (int64_t)W * C) : 0.f;
    const float e = (g.rin && g.bin) ? __ldg(tl + (int64_t)W * C + C) : 0.f;
    out[idx] = g.wx * g.wy * a + (1.f - g.wx) * g.wy * t + g.wx * (1.f - g.wy) * d +
               (1.f - g.wx) * (1.f - g.wy) * e;
  }
}

// ---- backward, C % 4 == 0 -------------------------------------------------------------
template <bool ONLY_GRID>
__global__ void __launch_bounds__(256)
warp_bwd_vec4(const float* __restrict__ img, const float* __restrict__ grid, const float* __restrict__ gout,
              float* __restrict__ gimg, float* __restrict__ ggrid, int B, int H, int W, int C, int Hg, int Wg) {
  const int64_t npix = (int64_t)B * Hg * Wg;
  const int sub = threadIdx.x & (LPP - 1);
  const int nch4 = C >> 2;
  // all lanes of a warp run the same number of iterations (npix is padded per warp) so that the
  // shuffles below are always executed by full warps
  const int64_t stride = (int64_t)gridDim.x * blockDim.x / LPP;
  const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPP;
  const int64_t warp_first = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / LPP;
  for (int64_t base = warp_first, pix = first; base < npix; base += stride, pix += stride) {
    const bool live = pix < npix;
    float d_tl = 0.f, d_tr = 0.f, d_bl = 0.f, d_br = 0.f;
    Geo g;
    g.wx = g.wy = 0.f;
    if (live) {
      const int xo = (int)(pix % Wg);
      const int yo = (int)((pix / Wg) % Hg);
      const int b = (int)(pix / ((int64_t)Wg * Hg));
      const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + pix);
      g = geometry(gxy.x, gxy.y, xo, yo, H, W);
      const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
      const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
      const int64_t a0 = (((int64_t)b * H + g.yi) * W + g.xi) * C;
      const float4* tl = reinterpret_cast<const float4*>(img + a0);
      const float4* tr = tl + nch4;
      const float4* bl = tl + (int64_t)W * nch4;
      const float4* br = bl + nch4;
      const float4* go = reinterpret_cast<const float4*>(gout + pix * C);
      float* gi = ONLY_GRID ? nullptr : gimg + a0;
      const bool both = g.rin && g.bin;
      for (int q = sub; q < nch4; q += LPP) {
        const float4 v = __ldg(go + q);
        {
          const float4 a = __ldg(tl + q);
          d_tl += a.x * v.x + a.y * v.y + a.z * v.z + a.w * v.w;
          if (!ONLY_GRID) red_add_v4(gi + 4 * q, w_tl * v.x, w_tl * v.y, w_tl * v.z, w_tl * v.w);
        }
        if (g.rin) {
          const float4 a = __ldg(tr + q);
          d_tr += a.x * v.x + a.y * v.y + a.z * v.z + a.w * v.w;
          if (!ONLY_GRID) red_add_v4(gi + C + 4 * q, w_tr * v.x, w_tr * v.y, w_tr * v.z, w_tr * v.w);
        }
        if (g.bin) {
          const float4 a = __ldg(bl + q);
          d_bl += a.x * v.x + a.y * v.y + a.z * v.z + a.w * v.w;
          if (!ONLY_GRID)
            red_add_v4(gi + (int64_t)W * C + 4 * q, w_bl * v.x, w_bl * v.y, w_bl * v.z, w_bl * v.w);
        }
        if (both) {
          const float4 a = __ldg(br + q);
          d_br += a.x * v.x + a.y * v.y + a.z * v.z + a.w * v.w;
          if (!ONLY_GRID)
            red_add_v4(gi + (int64_t)W * C + C + 4 * q, w_br * v.x, w_br * v.y, w_br * v.z, w_br * v.w);
        }
      }
    }
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) {
      d_tl += __shfl_xor_sync(0xffffffffu, d_tl, o);
      d_tr += __shfl_xor_sync(0xffffffffu, d_tr, o);
      d_bl += __shfl_xor_sync(0xffffffffu, d_bl, o);
      d_br += __shfl_xor_sync(0xffffffffu, d_br, o);
    }
    if (live && sub == 0) {
      float2 r;
      r.x = -g.wy * d_tl + g.wy * d_tr - (1.f - g.wy) * d_bl + (1.f - g.wy) * d_br;
      r.y = -g.wx * d_tl + g.wx * d_bl - (1.f - g.wx) * d_tr + (1.f - g.wx) * d_br;
      reinterpret_cast<float2*>(ggrid)[pix] = r;
    }
  }
}

// ---- backward, any C: one thread per pixel --------------------------------------------
template <bool ONLY_GRID, int CT>
__global__ void __launch_bounds__(256)
warp_bwd_scalar(const float* __restrict__ img, const float* __restrict__ grid, const float* __restrict__ gout,
                float* __restrict__ gimg, float* __restrict__ ggrid, int B, int H, int W, int Crt, int Hg,
                int Wg) {
  const int C = CT > 0 ? CT : Crt;
  const int64_t npix = (int64_t)B * Hg * Wg;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(pix % Wg);
    const int yo = (int)((pix / Wg) % Hg);
    const int b = (int)(pix / ((int64_t)Wg * Hg));
    const float2 gxy = __ldg(reinterpret_cast<const float2*>(grid) + pix);
    const Geo g = geometry(gxy.x, gxy.y, xo, yo, H, W);
    const float w_tl = g.wx * g.wy, w_tr = (1.f - g.wx) * g.wy;
    const float w_bl = g.wx * (1.f - g.wy), w_br = (1.f - g.wx) * (1.f - g.wy);
    const int64_t a0 = (((int64_t)b * H + g.yi) * W + g.xi) * C;
    const int64_t rowC = (int64_t)W * C;
    const bool both = g.rin && g.bin;
    float d_tl = 0.f, d_tr = 0.f, d_bl = 0.f, d_br = 0.f;
#pragma unroll
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
}

int check_args(const float* img, const float* grid, int B, int H, int W, int C, int Hg, int Wg) {
  if (!img || !grid) return fail(B2F_EINVAL, "warp: NULL img/grid");
  if (B < 0 || H <= 0 || W <= 0 || C <= 0 || Hg <= 0 || Wg <= 0)
    return fail(B2F_EINVAL, "warp: bad size B=%d H=%d W=%d C=%d Hg=%d Wg=%d", B, H, W, C, Hg, Wg);
  if (!aligned4(img)) return fail(B2F_EALIGN, "warp: img misaligned");
  if ((reinterpret_cast<uintptr_t>(grid) & 7u) != 0) return fail(B2F_EALIGN, "warp: grid must be 8-byte aligned");
  return B2F_OK;
}

int blocks_for(int64_t threads_needed) {
  int64_t blocks = (threads_needed + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
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
  const int64_t npix = (int64_t)B * Hg * Wg;
  if ((C & 3) == 0 && aligned16(img) && aligned16(out)) {
    warp_fwd_vec4<<<blocks_for(npix * LPP), 256, 0, st>>>(img, grid, out, B, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_fwd_vec4");
  } else if (C == 3) {
    warp_fwd_scalar<3><<<blocks_for(npix * 3), 256, 0, st>>>(img, grid, out, B, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_fwd_scalar<3>");
  } else {
    warp_fwd_scalar<0><<<blocks_for(npix * C), 256, 0, st>>>(img, grid, out, B, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_fwd_scalar<0>");
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
  const int64_t npix = (int64_t)B * Hg * Wg;
  const bool only = gradImg == nullptr;
  if ((C & 3) == 0 && aligned16(img) && aligned16(gradOut) && (only || aligned16(gradImg))) {
    // pad to whole warps: 4 pixels per warp
    const int64_t threads = ((npix + 3) / 4) * 32;
    if (only) warp_bwd_vec4<true><<<blocks_for(threads), 256, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, B, H, W, C, Hg, Wg);
    else warp_bwd_vec4<false><<<blocks_for(threads), 256, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, B, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_bwd_vec4");
  } else if (C == 3) {
    if (only) warp_bwd_scalar<true, 3><<<blocks_for(npix), 256, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, B, H, W, C, Hg, Wg);
    else warp_bwd_scalar<false, 3><<<blocks_for(npix), 256, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, B, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_bwd_scalar<3>");
  } else {
    if (only) warp_bwd_scalar<true, 0><<<blocks_for(npix), 256, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, B, H, W, C, Hg, Wg);
    else warp_bwd_scalar<false, 0><<<blocks_for(npix), 256, 0, st>>>(img, grid, gradOut, gradImg, gradGrid, B, H, W, C, Hg, Wg);
    B2F_CHECK_LAUNCH("warp_bwd_scalar<0>");
  }
  return B2F_OK;
}
