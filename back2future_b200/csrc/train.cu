// Training-only kernels of the conv trunk ("and their backward + Adam"): weight / bias gradient of the
// 3x3 convolution, the backward of the small layout ops, the gradient bookkeeping (axpy over strided channel slices,
// LeakyReLU derivative) and the Adam step (optim.adam, train.lua:485-486).
//
// Replaces, for the modules of models/pwc.lua: SpatialConvolution:accGradParameters, LeakyReLU:updateGradInput,
// SpatialUpSamplingBilinear / SpatialUpSamplingNearest / SpatialSoftMax :updateGradInput, nngraph's gradient
// accumulation at fan-out nodes, and optim.adam.
#include "tma.cuh"

#include <algorithm>

namespace b2f {
namespace {

// ---- weight gradient ---------------------------------------------------------------------------------------------
// gw[(ci * 9 + tap)][co] += sum_{b, y, x} gout[b, co, y, x] * x[b, ci, y S + ky - 1, x S + kx - 1]      (packed layout)
// gb[co]                 += sum_{b, y, x} gout[b, co, y, x]
//
// A GEMM with K = B * Ho * Wo (pixels): CTA = (8 input channels, 64 output channels, a share of the pixel tiles).
// Thread (co = tid % 32, ci = tid / 32) owns the 9 taps of (ci, co) and (ci, co + 32): 18 accumulators.  A 4 x 32 pixel
// tile of gout (64 planes) and the matching input patch (8 planes, + halo) are staged in shared memory with cp.async,
// two stages deep; per 4 pixels a thread reads its two gout float4 (conflict-free: the plane pitch is 4 mod 32 words)
// and, per tap row, 6 (stride 1) or 9 (stride 2) input values that are the same for the whole warp (broadcast) for
// 72 FMAs.  The partial sums leave with atomicAdd (co is the lane index: 128-byte coalesced reductions).  The input
// values of a row are read as one word + one or two aligned float4 (+ one word): 3 shared-memory loads instead of 6 / 9.
// Layers with at most 32 output channels (where the wide form idles half or more of its lanes) use the NARROW form:
// CTA = (16 input channels, 32 output channels), thread (co, ci) and (co, ci + 8).
namespace wg {
constexpr int TH = 4, TW = 32, THREADS = 256;
constexpr int GP = TW + 4;                 // gout row pitch (words)
constexpr int GPLANE = TH * GP + 4;        // plane pitch: 148 = 4 mod 32
// MODE 0 (wide):   CTA = 8 input x 64 output channels, thread (co, co + 32) x ci;
// MODE 1 (narrow, Cout <= 32: the pyramid's 16- and 32-channel layers, where half or three quarters of the wide form's
//                 lanes idle): CTA = 16 input x 32 output channels, thread co x (ci, ci + 8);
// MODE 2 (tiny, Cin <= 4 and Cout <= 16: the first convolution, 3 image channels -> 16): CTA = 4 input x 16 output
//                 channels x the 4 rows of the pixel tile, thread (co, ci, row).
template <int S, int MODE>
struct Cfg {
  static constexpr int KC = MODE == 2 ? 4 : (MODE == 1 ? 16 : 8), NC = MODE == 2 ? 16 : (MODE == 1 ? 32 : 64);
  static constexpr int XR = (TH - 1) * S + 3;          // input rows of a tile
  static constexpr int XW = (TW - 1) * S + 3;          // input columns
  static constexpr int XP = (XW + 3) / 4 * 4 + 4;      // row pitch, first column at word 3 so that column 1 is 16-byte aligned
  static constexpr int G_ELEMS = NC * GPLANE;
  static constexpr int X_ELEMS = KC * XR * XP;
  static constexpr int STAGE = G_ELEMS + X_ELEMS;
  static constexpr int SMEM_BYTES = 2 * STAGE * 4;
  static constexpr int NV = (XW - 1) / 4;              // whole 16-byte chunks of an input row after its first column
  static constexpr int RW = 1 + NV + ((XW - 1) % 4);   // copies per input row in the vector form (1 + 8 + 1 / 1 + 16)
  static_assert((XW - 1) % 4 <= 1, "one trailing column at most");
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = smem_u32(dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool valid) {
  const uint32_t d = smem_u32(dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// VEC: 16-byte staging copies of gout AND of the input rows (Wo % 4 == 0, W % 4 == 0, 16-byte aligned bases and batch
// strides): a row is its first column (the left halo), NV whole chunks and, for stride 1, one trailing column.
template <int S, bool VEC, int MODE>
__global__ void __launch_bounds__(THREADS, 2)
conv3x3_wgrad(const float* __restrict__ x, int64_t xbs, const float* __restrict__ gout, int64_t gbs,
              float* __restrict__ gw, float* __restrict__ gb, int B, int Cin, int H, int W, int Cout, int CoutP, int Ho,
              int Wo, int nco, int nsplit) {
  using cfg = Cfg<S, MODE>;
  constexpr int KC = cfg::KC, NC = cfg::NC;
  constexpr bool NARROW = MODE == 1, TINY = MODE == 2;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int co_l = TINY ? (tid & 15) : (tid & 31), ci_l = TINY ? ((tid >> 4) & 3) : (tid >> 5);
  const int ps = tid >> 6;                 // TINY: the tile row of this thread
  const int cblk = blockIdx.x;
  const int c0 = (cblk / nco) * KC, n0 = (cblk % nco) * NC;
  const int tiles_x = (Wo + TW - 1) / TW, tiles_y = (Ho + TH - 1) / TH;
  const int ntiles = B * tiles_y * tiles_x;
  const int t_lo = (int)((int64_t)ntiles * blockIdx.y / nsplit), t_hi = (int)((int64_t)ntiles * (blockIdx.y + 1) / nsplit);

  // acc[h]: WIDE h = output-channel half (co, co + 32); NARROW h = input-channel half (ci, ci + 8)
  float acc[2][9];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[h][t] = 0.f;
  float bsum[2] = {0.f, 0.f};

  auto stage = [&](int t, int s) {
    float* gs = smem + s * cfg::STAGE;
    float* xs = gs + cfg::G_ELEMS;
    const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
    const int x0 = tx * TW, y0 = ty * TH;
    const float* gbase = gout + (size_t)b * gbs;
    if (VEC) {
      // NC planes x 4 rows x 8 float4
      for (int i = tid; i < NC * TH * (TW / 4); i += THREADS) {
        const int q = i & 7, r = (i >> 3) & 3, c = i >> 5;
        const int yy = y0 + r, xx = x0 + 4 * q, n = n0 + c;
        const bool ok = n < Cout && yy < Ho && xx < Wo;      // Wo % 4 == 0: a float4 is inside or outside as a whole
        cp_async16(gs + c * GPLANE + r * GP + 4 * q, ok ? gbase + ((size_t)n * Ho + yy) * Wo + xx : gout, ok);
      }
    } else {
      for (int i = tid; i < NC * TH * TW; i += THREADS) {
        const int q = i & 31, r = (i >> 5) & 3, c = i >> 7;
        const int yy = y0 + r, xx = x0 + q, n = n0 + c;
        const bool ok = n < Cout && yy < Ho && xx < Wo;
        cp_async4(gs + c * GPLANE + r * GP + q, ok ? gbase + ((size_t)n * Ho + yy) * Wo + xx : gout, ok);
      }
    }
    const float* xbase = x + (size_t)b * xbs;
    const int xi0 = x0 * S - 1, yi0 = y0 * S - 1;
    if (VEC) {
      for (int i = tid; i < KC * cfg::XR * cfg::RW; i += THREADS) {
        const int it = i % cfg::RW, r = (i / cfg::RW) % cfg::XR, c = i / (cfg::RW * cfg::XR);
        const int yy = yi0 + r, ci = c0 + c;
        const bool rok = ci < Cin && yy >= 0 && yy < H;
        float* drow = xs + (c * cfg::XR + r) * cfg::XP + 3;
        const float* srow = xbase + ((size_t)ci * H + yy) * W;
        if (it >= 1 && it <= cfg::NV) {
          const int q = 1 + 4 * (it - 1), xx = xi0 + q;        // xx = x0 * S + 4 (it - 1): a multiple of 4
          const bool ok = rok && xx < W;                       // W % 4 == 0
          cp_async16(drow + q, ok ? srow + xx : x, ok);
        } else {
          const int q = it == 0 ? 0 : cfg::XW - 1, xx = xi0 + q;
          const bool ok = rok && xx >= 0 && xx < W;
          cp_async4(drow + q, ok ? srow + xx : x, ok);
        }
      }
    } else {
      for (int i = tid; i < KC * cfg::XR * cfg::XW; i += THREADS) {
        const int q = i % cfg::XW, r = (i / cfg::XW) % cfg::XR, c = i / (cfg::XW * cfg::XR);
        const int yy = yi0 + r, xx = xi0 + q, ci = c0 + c;
        const bool ok = ci < Cin && yy >= 0 && yy < H && xx >= 0 && xx < W;
        cp_async4(xs + (c * cfg::XR + r) * cfg::XP + 3 + q, ok ? xbase + ((size_t)ci * H + yy) * W + xx : x, ok);
      }
    }
    cp_commit();
  };

  if (t_lo < t_hi) stage(t_lo, 0);
  for (int t = t_lo; t < t_hi; ++t) {
    const int s = (t - t_lo) & 1;
    if (t + 1 < t_hi) {
      stage(t + 1, s ^ 1);
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    const float* gs = smem + s * cfg::STAGE;
    const float* xs = gs + cfg::G_ELEMS + ci_l * cfg::XR * cfg::XP + 3;
    const float* g0 = gs + co_l * GPLANE;
#pragma unroll
    for (int rr = 0; rr < (TINY ? 1 : TH); ++rr) {
      const int r = TINY ? ps : rr;
#pragma unroll 2
      for (int q = 0; q < TW / 4; ++q) {
        const float4 a = *reinterpret_cast<const float4*>(g0 + r * GP + 4 * q);
        float4 c = a;
        if (MODE == 0) c = *reinterpret_cast<const float4*>(g0 + 32 * GPLANE + r * GP + 4 * q);
        const float ga[4] = {a.x, a.y, a.z, a.w}, gc[4] = {c.x, c.y, c.z, c.w};
        if (ci_l == 0) {
          bsum[0] += (a.x + a.y) + (a.z + a.w);
          if (MODE == 0) bsum[1] += (c.x + c.y) + (c.z + c.w);
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          constexpr int NX = 3 * S + 3;      // 6 / 9 input values of this row: one word, then one or two aligned float4, [one word]
#pragma unroll
          for (int hx = 0; hx < (NARROW ? 2 : 1); ++hx) {
            const float* xr = xs + hx * 8 * cfg::XR * cfg::XP + (r * S + ky) * cfg::XP + 4 * q * S;
            float xv[NX];
            xv[0] = xr[0];
            {
              const float4 v = *reinterpret_cast<const float4*>(xr + 1);
              xv[1] = v.x; xv[2] = v.y; xv[3] = v.z; xv[4] = v.w;
            }
            if (S == 2) {
              const float4 v = *reinterpret_cast<const float4*>(xr + 5);
              xv[5] = v.x; xv[6] = v.y; xv[7] = v.z; xv[NX - 1] = v.w;
            } else {
              xv[NX - 1] = xr[NX - 1];
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                if (MODE != 0) {
                  acc[hx][ky * 3 + kx] = fmaf(ga[p], xv[p * S + kx], acc[hx][ky * 3 + kx]);
                } else {
                  acc[0][ky * 3 + kx] = fmaf(ga[p], xv[p * S + kx], acc[0][ky * 3 + kx]);
                  acc[1][ky * 3 + kx] = fmaf(gc[p], xv[p * S + kx], acc[1][ky * 3 + kx]);
                }
              }
          }
        }
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int h = 0; h < (TINY ? 1 : 2); ++h) {
    const int n = n0 + co_l + (MODE == 0 ? 32 * h : 0);
    const int ci = c0 + ci_l + (NARROW ? 8 * h : 0);
    if (n >= Cout) continue;
    if (ci < Cin) {
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(gw + (size_t)(ci * 9 + t) * CoutP + n, acc[h][t]);
    }
    if (gb && ci_l == 0 && c0 == 0 && (MODE == 0 || h == 0)) atomicAdd(gb + n, bsum[h]);
  }
}

}  // namespace wg

// ---- small backward ops --------------------------------------------------------------------------------------------
int ew_grid(int64_t total, int threads) {
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

// g *= (act > 0 ? 1 : slope), rows of `row` elements with independent strides
__global__ void leaky_backward_kernel(float* __restrict__ g, int64_t gs, const float* __restrict__ act, int64_t as, int64_t row,
                                      int64_t rows, float slope) {
  const int64_t total = rows * row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row, c = i % row;
    if (!(act[r * as + c] > 0.f)) g[r * gs + c] *= slope;
  }
}

// dst += alpha * src over rows with independent strides (nngraph's gradient accumulation at fan-out nodes; the
// level_weights * opt.* scaling of train.lua:421-468)
__global__ void axpy2d_kernel(float* __restrict__ dst, int64_t ds, const float* __restrict__ src, int64_t ss, int64_t row,
                              int64_t rows, float alpha) {
  const int64_t total = rows * row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row, c = i % row;
    dst[r * ds + c] = fmaf(alpha, src[r * ss + c], dst[r * ds + c]);
  }
}

// SpatialUpSamplingBilinear(2):updateGradInput as a gather: input row i receives from the output rows h2 with
// floor(r h2) in {i - 1, i}; the weights are recomputed with the forward's fp32 expressions.
__device__ __forceinline__ void up_taps(int n_in, int n_out, float ratio, int i, int* idx, float* wgt, int& n) {
  n = 0;
  // candidates: r h2 in (i - 1, i + 1)  ->  h2 in ((i - 1) / r, (i + 1) / r); scan a safe superset
  int lo = ratio > 0.f ? (int)floorf((float)(i - 1) / ratio) - 1 : 0;
  int hi = ratio > 0.f ? (int)ceilf((float)(i + 1) / ratio) + 1 : n_out - 1;
  lo = max(lo, 0);
  hi = min(hi, n_out - 1);
  for (int h2 = lo; h2 <= hi; ++h2) {
    const float h1r = __fmul_rn(ratio, (float)h2);
    const int h1 = (int)h1r;
    const int h1p = h1 < n_in - 1 ? 1 : 0;
    const float l1 = __fsub_rn(h1r, (float)h1), l0 = __fsub_rn(1.f, l1);
    float w = 0.f;
    if (h1 == i) w += l0;
    if (h1 + h1p == i) w += l1;
    if (h1 == i || h1 + h1p == i) {
      if (n < 6) { idx[n] = h2; wgt[n] = w; ++n; }
    }
  }
}
__global__ void upsample_bilinear2_backward_kernel(const float* __restrict__ go, float* __restrict__ gi, int B, int C, int H,
                                                   int W, float mul, int accumulate) {
  const int Ho = 2 * H, Wo = 2 * W;
  const float rh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float rw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xw = (int)(i % W), yh = (int)((i / W) % H);
    const int64_t pl = i / ((int64_t)W * H);
    int iy[6], ix[6], ny, nx;
    float wy[6], wx[6];
    up_taps(H, Ho, rh, yh, iy, wy, ny);
    up_taps(W, Wo, rw, xw, ix, wx, nx);
    const float* g = go + pl * (int64_t)Ho * Wo;
    float acc = 0.f;
    for (int a = 0; a < ny; ++a) {
      float row = 0.f;
      for (int c = 0; c < nx; ++c) row = fmaf(wx[c], __ldg(g + (int64_t)iy[a] * Wo + ix[c]), row);
      acc = fmaf(wy[a], row, acc);
    }
    acc *= mul;
    gi[i] = accumulate ? gi[i] + acc : acc;
  }
}

// SpatialUpSamplingNearest(scale):updateGradInput: sum of the scale x scale block
__global__ void upsample_nearest_backward_kernel(const float* __restrict__ go, float* __restrict__ gi, int64_t planes, int H,
                                                 int W, int scale) {
  const int64_t total = planes * H * W;
  const int Wo = W * scale;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xw = (int)(i % W), yh = (int)((i / W) % H);
    const int64_t pl = i / ((int64_t)W * H);
    const float* g = go + (pl * H * scale + (int64_t)yh * scale) * Wo + (int64_t)xw * scale;
    float acc = 0.f;
    for (int a = 0; a < scale; ++a)
      for (int c = 0; c < scale; ++c) acc += __ldg(g + (int64_t)a * Wo + c);
    gi[i] = acc;
  }
}

// out (planes, 2H, 2W): x at the even coordinates, zeros elsewhere.  The input gradient of a stride-2 convolution is the
// stride-1 input gradient of this dilated output gradient (gin[y, x] = sum_k z[y + 1 - ky, x + 1 - kx] w[k] with
// z[2 yo, 2 xo] = gout[yo, xo]), which runs on the TMA / FFMA2 kernel: 4x the multiply-adds of the direct form, at ~20x
// its rate (the direct kernel took 10.3 of the training step's 40 ms for 3 % of its arithmetic).
__global__ void zero_insert2x_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t planes, int H, int W) {
  const int Wo = 2 * W;
  const int64_t total = planes * 2 * H * (Wo / 4);      // one float4 (two source pixels) per thread
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % (Wo / 4)), Y = (int)((i / (Wo / 4)) % (2 * H));
    const int64_t pl = i / ((int64_t)(Wo / 4) * 2 * H);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((Y & 1) == 0) {
      const float2 s = *reinterpret_cast<const float2*>(x + (pl * H + (Y >> 1)) * W + 2 * q);
      v.x = s.x;
      v.z = s.y;
    }
    *reinterpret_cast<float4*>(out + (pl * 2 * H + Y) * Wo + 4 * q) = v;
  }
}

// SpatialSoftMax:updateGradInput: gi = s * (go - sum_c go_c s_c)
__global__ void softmax_channels_backward_kernel(const float* __restrict__ s, const float* __restrict__ go,
                                                 float* __restrict__ gi, int B, int C, int64_t hw) {
  const int64_t total = (int64_t)B * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / hw, p = i % hw;
    const int64_t o = b * C * hw + p;
    float dot = 0.f;
    for (int c = 0; c < C; ++c) dot = fmaf(go[o + c * hw], s[o + c * hw], dot);
    for (int c = 0; c < C; ++c) gi[o + c * hw] = s[o + c * hw] * (go[o + c * hw] - dot);
  }
}

// optim.adam (torch/optim adam.lua): m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g g; x -= step * m / (sqrt(v) + eps),
// step = lr * sqrt(1 - b2^t) / (1 - b1^t) computed by the host; weight decay adds wd * x to g first.
__global__ void adam_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float step, float b1, float b2, float eps, float wd) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (wd != 0.f) gi = fmaf(wd, x[i], gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    x[i] = x[i] - step * (mi / (sqrtf(vi) + eps));
  }
}

}  // namespace
}  // namespace b2f

using namespace b2f;

extern "C" int b2f_conv3x3_backward_weights(const float* x, int64_t x_batch_stride, const float* gout,
                                            int64_t gout_batch_stride, float* gw_packed, float* gbias, int B, int Cin,
                                            int H, int W, int Cout, int stride, b2f_stream_t stream) {
  if (!x || !gout || !gw_packed) return fail(B2F_EINVAL, "conv3x3_backward_weights: NULL x / gout / gw_packed");
  if (B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "conv3x3_backward_weights: bad size");
  if (stride != 1 && stride != 2) return fail(B2F_EUNSUPPORTED, "conv3x3_backward_weights: stride %d", stride);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const int64_t xbs = x_batch_stride ? x_batch_stride : (int64_t)Cin * H * W;
  const int64_t gbs = gout_batch_stride ? gout_batch_stride : (int64_t)Cout * Ho * Wo;
  if (xbs < (int64_t)Cin * H * W || gbs < (int64_t)Cout * Ho * Wo)
    return fail(B2F_EINVAL, "conv3x3_backward_weights: batch stride smaller than one item");
  if (!aligned4(x) || !aligned4(gout) || !aligned4(gw_packed) || (gbias && !aligned4(gbias)))
    return fail(B2F_EALIGN, "conv3x3_backward_weights: misaligned pointer");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int CoutP = (Cout + 63) / 64 * 64;
  const int mode = (Cin <= 4 && Cout <= 16) ? 2 : ((Cout <= 32 && Cin > 8) ? 1 : 0);
  const int KC = mode == 2 ? 4 : (mode == 1 ? 16 : 8), NC = mode == 2 ? 16 : (mode == 1 ? 32 : 64);
  const int nci = (Cin + KC - 1) / KC, nco = (Cout + NC - 1) / NC;
  const int ntiles = B * ((Ho + wg::TH - 1) / wg::TH) * ((Wo + wg::TW - 1) / wg::TW);
  // enough CTAs for ~3 waves of 2 per SM, at least 4 tiles per CTA so the two-stage pipeline has something to overlap
  int nsplit = (num_sms() * 6 + nci * nco - 1) / (nci * nco);
  nsplit = std::max(1, std::min(nsplit, std::max(1, ntiles / 4)));
  if (nsplit > 65535) nsplit = 65535;
  const bool vec = (Wo % 4) == 0 && aligned16(gout) && gbs % 4 == 0 && (W % 4) == 0 && aligned16(x) && xbs % 4 == 0;
  dim3 grid(nci * nco, nsplit);
#define B2F_WG(S, V, NRW)                                                                                             \
  do {                                                                                                                \
    auto kern = wg::conv3x3_wgrad<S, V, NRW>;                                                                         \
    static thread_local int attr_dev = -1;                                                                            \
    int dev = 0;                                                                                                      \
    B2F_CUDA_TRY(cudaGetDevice(&dev));                                                                                \
    if (attr_dev != dev) {                                                                                            \
      B2F_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::Cfg<S, NRW>::SMEM_BYTES)); \
      attr_dev = dev;                                                                                                 \
    }                                                                                                                 \
    kern<<<grid, wg::THREADS, wg::Cfg<S, NRW>::SMEM_BYTES, st>>>(x, xbs, gout, gbs, gw_packed, gbias, B, Cin, H, W, Cout, \
                                                                  CoutP, Ho, Wo, nco, nsplit);                        \
  } while (0)
#define B2F_WG2(S, V)                                                                                                 \
  do {                                                                                                                \
    if (mode == 2) B2F_WG(S, V, 2); else if (mode == 1) B2F_WG(S, V, 1); else B2F_WG(S, V, 0);                         \
  } while (0)
  if (stride == 1) {
    if (vec) B2F_WG2(1, true); else B2F_WG2(1, false);
  } else {
    if (vec) B2F_WG2(2, true); else B2F_WG2(2, false);
  }
#undef B2F_WG2
#undef B2F_WG
  B2F_CHECK_LAUNCH("conv3x3_wgrad");
  return B2F_OK;
}

extern "C" int b2f_leaky_relu_backward(float* grad, int64_t grad_row_stride, const float* act, int64_t act_row_stride,
                                       int64_t row_elems, int64_t rows, float slope, b2f_stream_t stream) {
  if (rows < 0 || row_elems < 0) return fail(B2F_EINVAL, "leaky_relu_backward: negative extent");
  if (!rows || !row_elems) return B2F_OK;
  if (!grad || !act) return fail(B2F_EINVAL, "leaky_relu_backward: NULL pointer");
  if (grad_row_stride < row_elems || act_row_stride < row_elems) return fail(B2F_EINVAL, "leaky_relu_backward: bad stride");
  leaky_backward_kernel<<<ew_grid(rows * row_elems, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      grad, grad_row_stride, act, act_row_stride, row_elems, rows, slope);
  B2F_CHECK_LAUNCH("leaky_backward_kernel");
  return B2F_OK;
}

extern "C" int b2f_axpy2d(float* dst, int64_t dst_row_stride, const float* src, int64_t src_row_stride, int64_t row_elems,
                          int64_t rows, float alpha, b2f_stream_t stream) {
  if (rows < 0 || row_elems < 0) return fail(B2F_EINVAL, "axpy2d: negative extent");
  if (!rows || !row_elems) return B2F_OK;
  if (!dst || !src) return fail(B2F_EINVAL, "axpy2d: NULL pointer");
  if (dst_row_stride < row_elems || src_row_stride < row_elems) return fail(B2F_EINVAL, "axpy2d: bad stride");
  axpy2d_kernel<<<ew_grid(rows * row_elems, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dst, dst_row_stride, src, src_row_stride, row_elems, rows, alpha);
  B2F_CHECK_LAUNCH("axpy2d_kernel");
  return B2F_OK;
}

extern "C" int b2f_upsample_bilinear2x_backward(const float* grad_out, float* grad_in, int B, int C, int H, int W, float mul,
                                                int accumulate, b2f_stream_t stream) {
  if (!grad_out || !grad_in || B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "upsample_bilinear2x_backward: bad argument");
  if (B == 0) return B2F_OK;
  upsample_bilinear2_backward_kernel<<<ew_grid((int64_t)B * C * H * W, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      grad_out, grad_in, B, C, H, W, mul, accumulate);
  B2F_CHECK_LAUNCH("upsample_bilinear2_backward_kernel");
  return B2F_OK;
}

extern "C" int b2f_zero_insert2x(const float* x, float* out, int64_t planes, int H, int W, b2f_stream_t stream) {
  if (!x || !out || planes < 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "zero_insert2x: bad argument");
  if ((W & 1) || (reinterpret_cast<uintptr_t>(x) & 7u) || !aligned16(out))
    return fail(B2F_EUNSUPPORTED, "zero_insert2x: odd width or misaligned buffers");
  if (planes == 0) return B2F_OK;
  zero_insert2x_kernel<<<ew_grid(planes * 2 * H * (W / 2), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, planes, H, W);
  B2F_CHECK_LAUNCH("zero_insert2x_kernel");
  return B2F_OK;
}

extern "C" int b2f_upsample_nearest_backward(const float* grad_out, float* grad_in, int B, int C, int H, int W, int scale,
                                             b2f_stream_t stream) {
  if (!grad_out || !grad_in || B < 0 || C <= 0 || H <= 0 || W <= 0 || scale < 1) return fail(B2F_EINVAL, "upsample_nearest_backward: bad argument");
  if (B == 0) return B2F_OK;
  upsample_nearest_backward_kernel<<<ew_grid((int64_t)B * C * H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      grad_out, grad_in, (int64_t)B * C, H, W, scale);
  B2F_CHECK_LAUNCH("upsample_nearest_backward_kernel");
  return B2F_OK;
}

extern "C" int b2f_softmax_channels_backward(const float* softmax_out, const float* grad_out, float* grad_in, int B, int C,
                                             int H, int W, b2f_stream_t stream) {
  if (!softmax_out || !grad_out || !grad_in || B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "softmax_channels_backward: bad argument");
  if (B == 0) return B2F_OK;
  softmax_channels_backward_kernel<<<ew_grid((int64_t)B * H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      softmax_out, grad_out, grad_in, B, C, (int64_t)H * W);
  B2F_CHECK_LAUNCH("softmax_channels_backward_kernel");
  return B2F_OK;
}

extern "C" int b2f_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int64_t t, b2f_stream_t stream) {
  if (n < 0 || t < 1) return fail(B2F_EINVAL, "adam_step: n >= 0 and t >= 1 required (t = %lld)", (long long)t);
  if (!n) return B2F_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq) return fail(B2F_EINVAL, "adam_step: NULL pointer");
  // adam.lua: biasCorrection1 = 1 - beta1^t, biasCorrection2 = 1 - beta2^t, stepSize = lr * sqrt(bc2) / bc1 (doubles)
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  const float step = (float)((double)lr * sqrt(bc2) / bc1);
  adam_kernel<<<ew_grid(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, step,
                                                                                  beta1, beta2, eps, weight_decay);
  B2F_CHECK_LAUNCH("adam_kernel");
  return B2F_OK;
}
