// Conv trunk of the PWC network: 3x3 convolution (+ bias + LeakyReLU) and the small layout
// ops around it, for sm_100a.  Replaces cudnn.SpatialConvolution(nIn, nOut, 3, 3, s, s, 1, 1) + nn.LeakyReLU(0.2)
// (models/pwc.lua:58-65 convUnit, :76-85 decoder), nn.SpatialAveragePooling(2,2,2,2) (:153),
// nn.SpatialUpSamplingBilinear(2) (:359), nn.SpatialUpSamplingNearest(2) (:308-316), nn.SpatialSoftMax (:305).
//
// The convolution is an implicit GEMM on the fp32 FMA pipe (M = pixels, N = output channels, K = 9 Cin), fp32
// parity at 1e-4 through a 6-layer decoder and 5 pyramid levels leaves no room for a single-pass TF32/BF16 product:
//   * CTA = 8 warps (lane 0 of warp 0 also issues the TMA loads).  A warp owns 8 output channels x (8 rows x 32 columns);
//     a thread owns 8 consecutive pixels x 8 channels = 64 accumulators held as 32 packed pairs (FFMA2: the pair
//     operand is two neighbouring output channels of the weight row, the other operand the input pixel broadcast).
//   * Per pipeline stage the producer lands KC input channels of the input tile + halo with ONE TMA box -- the box
//     starts one row / four columns outside the tile and out-of-bounds elements read as zero, which IS the
//     convolution's zero padding -- and the KC x 9 x NT slice of the packed weights with a second box.  Full/empty
//     mbarrier ring, 3 stages.
//   * Weights are kept in the packed layout [Cin * 9][CoutP] (CoutP = Cout rounded up to 64): a thread's 8 output channels
//     of one tap are 32 contiguous bytes that every lane of the warp reads from the same address (broadcast).
//     b2f_conv3x3_pack_weights converts from / to Torch's (Cout, Cin, 3, 3).
//   * Input and output may be channel slices of wider buffers (base pointer + batch stride): the cost volumes, the
//     reference features and the up-sampled flow are written straight into the decoder's joined input
//     (nn.JoinTable, pwc.lua:267, 298-305, 334, disappears); `out2` stores a second copy for a second consumer.
#include "tma.cuh"

#include <algorithm>

namespace b2f {
namespace {

namespace cv3 {
constexpr int TH = 8, TW = 32;   // output pixels of one compute warp
constexpr int STAGES = 3;
constexpr int NCW = 8;           // compute warps
constexpr int THREADS = NCW * 32;   // no producer warp: nine warps on four schedulers cap the registers at 96

template <int NWN_, int S_, int KC_>
struct Cfg {
  static constexpr int NWN = NWN_, S = S_, KC = KC_;
  static constexpr int NWP = NCW / NWN;           // pixel blocks stacked in y
  static constexpr int NT = NWN * 8;              // output channels per CTA
  static constexpr int OH = TH * NWP;             // output rows per CTA
  static constexpr int IR = (OH - 1) * S + 3;     // input rows of the tile
  // input columns: the box starts 4 columns left of the tile (16-byte aligned start); row pitch is an odd multiple
  // of 16 bytes so that the rows a quarter-warp reads sit on disjoint bank groups
  static constexpr int IW = S == 1 ? 44 : 76;
  static constexpr int IN_BYTES = KC * IR * IW * 4;
  static constexpr int W_BYTES = KC * 9 * NT * 4;
  static constexpr int IN_PAD = (IN_BYTES + 127) / 128 * 128;
  static constexpr int STAGE_BYTES = IN_PAD + W_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;
  static constexpr int CTAS_PER_SM = (2 * (SMEM_BYTES + 1024) <= 228 * 1024) ? 2 : 1;
  static_assert(W_BYTES % 128 == 0, "TMA destination alignment");
  static_assert((TW - 1) * S + 3 + 3 <= IW, "tile + halo must fit the box");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct Args {
  const float* bias;       // [Cout] or NULL
  float* out;              // (B, Cout, Ho, Wo), batch stride obs
  float* out2;             // optional second destination, batch stride obs2
  int64_t obs, obs2;
  int Cin, Cout, Ho, Wo, ntn;
  float slope;             // LeakyReLU negative slope; 1 = no activation
  // backward-data use (b2f_conv3x3_backward_data): the activation derivative comes from `mask` (the forward output of
  // the layer whose input gradient this is: factor 1 where mask > 0, `slope` elsewhere) instead of the result's own
  // sign, and the result may be added to what `out` already holds
  const float* mask;
  int64_t mbs;
  int accumulate;
};

template <int NWN, int S, int KC>
__global__ void __launch_bounds__(THREADS, (Cfg<NWN, S, KC>::CTAS_PER_SM))
conv3x3_tma(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w, const Args a) {
  using cfg = Cfg<NWN, S, KC>;
  constexpr int NT = cfg::NT, IW = cfg::IW, IR = cfg::IR;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * cfg::OH;
  const int nt = blockIdx.z % a.ntn, b = blockIdx.z / a.ntn;
  const int n0 = nt * NT;
  const int niter = (a.Cin + KC - 1) / KC;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NCW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // lane 0 of warp 0 is the TMA producer, two stages ahead of the arithmetic (a tenth warp would put five warps on
  // two of the four schedulers and cap every thread at 96 registers)
  auto issue = [&](int it) {
    const int s = it % STAGES;
    uint8_t* st = smem + s * cfg::STAGE_BYTES;
    mbar_arrive_expect_tx(&full[s], cfg::IN_BYTES + cfg::W_BYTES);
    tma_load_4d(st, &tm_in, x0 * S - 4, y0 * S - 1, it * KC, b, &full[s]);
    tma_load_2d(st + cfg::IN_PAD, &tm_w, n0, it * KC * 9, &full[s]);
  };
  if (threadIdx.x == 0)
    for (int it = 0; it < STAGES - 1 && it < niter; ++it) issue(it);

  // ---- consumers ----
  const int wn = warp % NWN, pb = warp / NWN;
  const int r = lane >> 2, cgx = lane & 3;
  f32x2 acc[8][4];   // [pixel][channel pair]
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[p][m] = 0ull;

  // first input element a thread needs in its row: the left neighbour of its first pixel's centre tap
  const uint32_t xoff = (uint32_t)(((pb * TH + r) * S) * IW + (8 * S) * cgx + 3) * 4u;
  const uint32_t woff = (uint32_t)cfg::IN_PAD + (uint32_t)(wn * 8) * 4u;
  constexpr int NX = 7 * S + 3;   // input values of one row a thread touches

#pragma unroll 1
  for (int it = 0; it < niter; ++it) {
    const int s = it % STAGES;
    if (threadIdx.x == 0) {
      const int nxt = it + STAGES - 1;
      if (nxt < niter) {
        if (nxt >= STAGES) mbar_wait(&empty[nxt % STAGES], ((nxt / STAGES) - 1) & 1);
        issue(nxt);
      }
    }
    __syncwarp();
    mbar_wait(&full[s], (it / STAGES) & 1);
    const uint32_t sb = smem_u32(smem + s * cfg::STAGE_BYTES);
#pragma unroll 1
    for (int ci = 0; ci < KC; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const uint32_t xa = sb + xoff + (uint32_t)((ci * IR + ky) * IW) * 4u;
        float xin[NX + 1];
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xin[0]) : "r"(xa));
#pragma unroll
        for (int q = 0; q < (NX - 1) / 4; ++q) {
          const float4 v = lds128(xa + 4u + 16u * q);
          xin[1 + 4 * q] = v.x; xin[2 + 4 * q] = v.y; xin[3 + 4 * q] = v.z; xin[4 + 4 * q] = v.w;
        }
        if (S == 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xin[9]) : "r"(xa + 36u));
        const uint32_t wa = sb + woff + (uint32_t)((ci * 9 + ky * 3) * NT) * 4u;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          f32x2 w2[4];
          lds128_pairs(wa + (uint32_t)(kx * NT) * 4u, w2[0], w2[1]);
          lds128_pairs(wa + (uint32_t)(kx * NT) * 4u + 16u, w2[2], w2[3]);
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const f32x2 sx = pack2(xin[S * p + kx], xin[S * p + kx]);
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[p][m] = fma2(sx, w2[m], acc[p][m]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ---- epilogue: bias, LeakyReLU, stores (up to two destinations) ----
  const int y = y0 + pb * TH + r;
  const int x = x0 + 8 * cgx;
  if (y >= a.Ho || x >= a.Wo) return;
  const size_t hw = (size_t)a.Ho * a.Wo;
  const bool vec = ((a.Wo & 3) == 0);   // rows are then 16-byte aligned (the host checks base and batch strides)
#pragma unroll
  for (int m = 0; m < 4; ++m) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + wn * 8 + 2 * m + h;
      if (n >= a.Cout) continue;
      const float bv = a.bias ? __ldg(a.bias + n) : 0.f;
      const size_t off = (size_t)n * hw + (size_t)y * a.Wo + x;
      float v[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        float lo, hi;
        unpack2(acc[p][m], lo, hi);
        float t = (h ? hi : lo) + bv;
        if (a.mask) {
          const float mk = (x + p < a.Wo) ? __ldg(a.mask + (size_t)b * a.mbs + off + p) : 1.f;
          t = mk > 0.f ? t : t * a.slope;
        } else {
          t = t > 0.f ? t : t * a.slope;
        }
        if (a.accumulate && x + p < a.Wo) t += a.out[(size_t)b * a.obs + off + p];
        v[p] = t;
      }
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        float* dst = d == 0 ? a.out : a.out2;
        if (!dst) continue;
        dst += (size_t)b * (d == 0 ? a.obs : a.obs2) + off;
        if (vec && x + 8 <= a.Wo) {
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
          for (int p = 0; p < 8; ++p)
            if (x + p < a.Wo) dst[p] = v[p];
        }
      }
    }
  }
}

}  // namespace cv3

// Direct form for shapes the TMA path cannot take (row pitch not a multiple of 16 bytes: the two coarsest levels of
// a 1216-wide input are 38 and 19 columns wide).  One thread per output element; a few hundred pixels at most.
__global__ void conv3x3_generic(const float* __restrict__ x, int64_t xbs, const float* __restrict__ wp,
                                const float* __restrict__ bias, float* __restrict__ out, int64_t obs,
                                float* __restrict__ out2, int64_t obs2, int B, int Cin, int H, int W, int Cout,
                                int CoutP, int Ho, int Wo, int S, float slope) {
  const int64_t total = (int64_t)B * Cout * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho);
    const int n = (int)((i / ((int64_t)Wo * Ho)) % Cout), b = (int)(i / ((int64_t)Wo * Ho * Cout));
    float acc = 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xc = x + (size_t)b * xbs + (size_t)ci * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yi = yo * S - 1 + ky;
        if (yi < 0 || yi >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xi = xo * S - 1 + kx;
          if (xi < 0 || xi >= W) continue;
          acc = fmaf(__ldg(xc + (size_t)yi * W + xi), __ldg(wp + (size_t)(ci * 9 + ky * 3 + kx) * CoutP + n), acc);
        }
      }
    }
    float t = acc + (bias ? __ldg(bias + n) : 0.f);
    t = t > 0.f ? t : t * slope;
    const size_t off = (size_t)n * Ho * Wo + (size_t)yo * Wo + xo;
    out[(size_t)b * obs + off] = t;
    if (out2) out2[(size_t)b * obs2 + off] = t;
  }
}

// (Cout, Cin, 3, 3) <-> [Cin * 9][CoutP]
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin, int CoutP,
                                    int unpack) {
  const int total = Cin * 9 * CoutP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % CoutP, k = (i / CoutP) % 9, ci = i / (CoutP * 9);
    if (unpack) {
      if (n < Cout) const_cast<float*>(w)[((size_t)n * Cin + ci) * 9 + k] = wp[i];
    } else {
      wp[i] = n < Cout ? w[((size_t)n * Cin + ci) * 9 + k] : 0.f;
    }
  }
}

// ---- small layout ops ------------------------------------------------------------------------------------------
// nn.SpatialAveragePooling(2, 2, 2, 2): out[y, x] = mean of the 2 x 2 block (floor mode: a trailing odd row / column
// is dropped); THNN divides the SUM by 4 (count_include_pad, no padding here).
__global__ void avgpool2_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t planes, int H, int W,
                                int Ho, int Wo) {
  const int64_t total = planes * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho);
    const int64_t pl = i / ((int64_t)Wo * Ho);
    const float* p = x + pl * (int64_t)H * W + (int64_t)(2 * yo) * W + 2 * xo;
    const float2 a = *reinterpret_cast<const float2*>(p);
    const float2 c = *reinterpret_cast<const float2*>(p + W);
    out[i] = (((a.x + a.y) + c.x) + c.y) / 4.f;   // THNN's accumulation order: row-major over the window
  }
}

struct UpDst {
  float* p[3];
  int64_t bs[3];
};

// nn.SpatialUpSamplingBilinear(2): THNN/THCUNN map dst -> src with the align-corners ratio (in - 1) / (out - 1):
// h1r = rheight * h2; h1 = (int)h1r; h1p = (h1 < H - 1); lambda1 = h1r - h1; lambda0 = 1 - lambda1 (fp32),
// out = l0h * (l0w * a + l1w * b) + l1h * (l0w * c + l1w * d).
__global__ void upsample_bilinear2_kernel(const float* __restrict__ x, int64_t xbs, UpDst dst, int B, int C, int H, int W,
                                          float mul) {
  const int Ho = 2 * H, Wo = 2 * W;
  const float rh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float rw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const int64_t total = (int64_t)B * C * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w2 = (int)(i % Wo), h2 = (int)((i / Wo) % Ho);
    const int c = (int)((i / ((int64_t)Wo * Ho)) % C), b = (int)(i / ((int64_t)Wo * Ho * C));
    const float h1r = __fmul_rn(rh, (float)h2);
    const int h1 = (int)h1r;
    const int h1p = h1 < H - 1 ? 1 : 0;
    const float h1l = __fsub_rn(h1r, (float)h1), h0l = __fsub_rn(1.f, h1l);
    const float w1r = __fmul_rn(rw, (float)w2);
    const int w1 = (int)w1r;
    const int w1p = w1 < W - 1 ? 1 : 0;
    const float w1l = __fsub_rn(w1r, (float)w1), w0l = __fsub_rn(1.f, w1l);
    const float* p = x + (size_t)b * xbs + ((size_t)c * H + h1) * W + w1;
    const float va = __ldg(p), vb = __ldg(p + w1p), vc = __ldg(p + h1p * W), vd = __ldg(p + h1p * W + w1p);
    const float top = __fadd_rn(__fmul_rn(w0l, va), __fmul_rn(w1l, vb));
    const float bot = __fadd_rn(__fmul_rn(w0l, vc), __fmul_rn(w1l, vd));
    float v = __fadd_rn(__fmul_rn(h0l, top), __fmul_rn(h1l, bot));
    if (mul != 1.f) v *= mul;
    const size_t off = ((size_t)c * Ho + h2) * Wo + w2;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (dst.p[d]) dst.p[d][(size_t)b * dst.bs[d] + off] = v;
  }
}

// nn.SpatialUpSamplingNearest(scale): out[y, x] = in[y / scale, x / scale]
__global__ void upsample_nearest_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t planes, int H, int W,
                                        int scale) {
  const int Ho = H * scale, Wo = W * scale;
  const int64_t total = planes * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho);
    const int64_t pl = i / ((int64_t)Wo * Ho);
    out[i] = __ldg(x + (pl * H + yo / scale) * W + xo / scale);
  }
}

// nn.SpatialSoftMax on (B, C, h, w): softmax over the channel dimension, per pixel, max-subtracted like THNN
__global__ void softmax_channels_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int64_t hw) {
  const int64_t total = (int64_t)B * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / hw, p = i % hw;
    const float* xp = x + b * C * hw + p;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, xp[(int64_t)c * hw]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(xp[(int64_t)c * hw] - mx);
    float* op = out + b * C * hw + p;
    for (int c = 0; c < C; ++c) op[(int64_t)c * hw] = expf(xp[(int64_t)c * hw] - mx) / sum;
  }
}

int ew_grid(int64_t total, int threads) {
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

template <int NWN, int S, int KC>
int launch_conv(const float* x, int64_t xbs, const float* wp, int CoutP, const cv3::Args& a, int B, int Cin, int H, int W,
                cudaStream_t st) {
  using cfg = cv3::Cfg<NWN, S, KC>;
  CUtensorMap tin, tw;
  const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)Cin, (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)W, (uint64_t)W * H, (uint64_t)xbs};
  const uint32_t box[4] = {(uint32_t)cfg::IW, (uint32_t)cfg::IR, (uint32_t)KC, 1};
  int rc = make_tmap4(&tin, x, dims, str, box);
  if (rc) return rc;
  if ((rc = make_tmap2(&tw, wp, (uint64_t)CoutP, (uint64_t)Cin * 9, (uint64_t)CoutP, (uint32_t)cfg::NT, (uint32_t)(KC * 9)))) return rc;
  auto kern = cv3::conv3x3_tma<NWN, S, KC>;
  static thread_local int attr_dev = -1;
  int dev = 0;
  B2F_CUDA_TRY(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    B2F_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    attr_dev = dev;
  }
  cv3::Args aa = a;
  aa.ntn = (a.Cout + cfg::NT - 1) / cfg::NT;
  const int64_t gz = (int64_t)B * aa.ntn;
  if (gz > 65535) return fail(B2F_EINVAL, "conv3x3: B * channel tiles = %lld exceeds the grid limit", (long long)gz);
  dim3 grid((a.Wo + cv3::TW - 1) / cv3::TW, (a.Ho + cfg::OH - 1) / cfg::OH, (unsigned)gz);
  kern<<<grid, cv3::THREADS, cfg::SMEM_BYTES, st>>>(tin, tw, aa);
  B2F_CHECK_LAUNCH("conv3x3_tma");
  return B2F_OK;
}

}  // namespace
}  // namespace b2f

using namespace b2f;

extern "C" int64_t b2f_conv3x3_packed_floats(int Cin, int Cout) {
  if (Cin <= 0 || Cout <= 0) return 0;
  return (int64_t)Cin * 9 * ((Cout + 63) / 64 * 64);
}

extern "C" int b2f_conv3x3_pack_weights(const float* w, float* packed, int Cout, int Cin, int unpack, b2f_stream_t stream) {
  if (!w || !packed || Cout <= 0 || Cin <= 0) return fail(B2F_EINVAL, "conv3x3_pack_weights: bad argument");
  const int CoutP = (Cout + 63) / 64 * 64;
  pack_weights_kernel<<<ew_grid((int64_t)Cin * 9 * CoutP, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, packed, Cout, Cin, CoutP, unpack);
  B2F_CHECK_LAUNCH("pack_weights_kernel");
  return B2F_OK;
}

extern "C" int b2f_conv3x3_forward(const float* x, int64_t x_batch_stride, const float* w_packed, const float* bias,
                                   float* out, int64_t out_batch_stride, float* out2, int64_t out2_batch_stride,
                                   int B, int Cin, int H, int W, int Cout, int stride, float leaky_slope,
                                   b2f_stream_t stream) {
  if (!x || !w_packed || !out) return fail(B2F_EINVAL, "conv3x3_forward: NULL x / w_packed / out");
  if (B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "conv3x3_forward: bad size B=%d Cin=%d Cout=%d H=%d W=%d", B, Cin, Cout, H, W);
  if (stride != 1 && stride != 2) return fail(B2F_EUNSUPPORTED, "conv3x3_forward: stride %d (the model uses 1 and 2)", stride);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const int64_t xbs = x_batch_stride ? x_batch_stride : (int64_t)Cin * H * W;
  const int64_t obs = out_batch_stride ? out_batch_stride : (int64_t)Cout * Ho * Wo;
  const int64_t obs2 = out2_batch_stride ? out2_batch_stride : (int64_t)Cout * Ho * Wo;
  if (xbs < (int64_t)Cin * H * W || obs < (int64_t)Cout * Ho * Wo || (out2 && obs2 < (int64_t)Cout * Ho * Wo))
    return fail(B2F_EINVAL, "conv3x3_forward: batch stride smaller than one item");
  if (!aligned4(x) || !aligned4(out) || !aligned4(w_packed) || (out2 && !aligned4(out2)))
    return fail(B2F_EALIGN, "conv3x3_forward: misaligned pointer");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int CoutP = (Cout + 63) / 64 * 64;
  cv3::Args a{bias, out, out2, obs, obs2, Cin, Cout, Ho, Wo, 1, leaky_slope, nullptr, 0, 0};

  const bool vec_ok = (Wo % 4) != 0 || (aligned16(out) && obs % 4 == 0 && (!out2 || (aligned16(out2) && obs2 % 4 == 0)));
  const bool tma_ok = (W % 4) == 0 && aligned16(x) && xbs % 4 == 0 && aligned16(w_packed) && vec_ok &&
                      get_encode_fn() != nullptr;
  if (tma_ok) {
    // channel tile: 64 output channels per CTA when that wastes nothing, else 32 (Cout = 96, 32, 16, 2)
    // ... and 16 for the first pyramid level and the heads (Cout = 16, 2): half of a 32-channel tile's arithmetic
    // would be padding (two channel warps x four pixel blocks: 32 output rows per CTA, fewer input channels per stage)
    const bool n64 = Cout % 64 == 0 || Cout > 96;
    if (Cout <= 16)
      return stride == 1 ? launch_conv<2, 1, 4>(x, xbs, w_packed, CoutP, a, B, Cin, H, W, st)
                         : launch_conv<2, 2, 2>(x, xbs, w_packed, CoutP, a, B, Cin, H, W, st);
    if (stride == 1)
      return n64 ? launch_conv<8, 1, 8>(x, xbs, w_packed, CoutP, a, B, Cin, H, W, st)
                 : launch_conv<4, 1, 8>(x, xbs, w_packed, CoutP, a, B, Cin, H, W, st);
    return n64 ? launch_conv<8, 2, 4>(x, xbs, w_packed, CoutP, a, B, Cin, H, W, st)
               : launch_conv<4, 2, 4>(x, xbs, w_packed, CoutP, a, B, Cin, H, W, st);
  }
  conv3x3_generic<<<ew_grid((int64_t)B * Cout * Ho * Wo, 128), 128, 0, st>>>(x, xbs, w_packed, bias, out, obs, out2, obs2,
                                                                            B, Cin, H, W, Cout, CoutP, Ho, Wo, stride,
                                                                            leaky_slope);
  B2F_CHECK_LAUNCH("conv3x3_generic");
  return B2F_OK;
}

// ---- backward-data ----------------------------------------------------------------------------------------------
// gin[ci, y, x] = sum_{co, ky, kx} gout[co, (y + 1 - ky) / s, (x + 1 - kx) / s] * w[co, ci, ky, kx]   (divisible terms)
// Stride 1 is the forward kernel again on the transposed, tap-flipped weights wT[(co * 9 + 8 - tap)][ci]
// (b2f_conv3x3_transpose_packed); stride 2 (the six down-sampling convolutions of the feature pyramid, 3 % of the
// network's arithmetic) and widths the TMA path cannot take use the direct kernel below on the same wT.
namespace b2f {
namespace {
__global__ void conv3x3_dgrad_generic(const float* __restrict__ gout, int64_t gbs, const float* __restrict__ wt,
                                      const float* __restrict__ act, int64_t abs_, float* __restrict__ gin, int64_t ibs,
                                      int accumulate, int B, int Cout, int Ho, int Wo, int Cin, int CinP, int H, int W,
                                      int S, float slope) {
  const int64_t total = (int64_t)B * Cin * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int ci = (int)((i / ((int64_t)W * H)) % Cin), b = (int)(i / ((int64_t)W * H * Cin));
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || ty % S) continue;
      const int yo = ty / S;
      if (yo >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || tx % S) continue;
        const int xo = tx / S;
        if (xo >= Wo) continue;
        const float* g = gout + (size_t)b * gbs + (size_t)yo * Wo + xo;
        const float* w = wt + (size_t)(8 - (ky * 3 + kx)) * CinP + ci;
        for (int co = 0; co < Cout; ++co)
          acc = fmaf(__ldg(g + (size_t)co * Ho * Wo), __ldg(w + (size_t)co * 9 * CinP), acc);
      }
    }
    const size_t off = ((size_t)ci * H + y) * W + x;
    if (act) acc = __ldg(act + (size_t)b * abs_ + off) > 0.f ? acc : acc * slope;
    float* d = gin + (size_t)b * ibs + off;
    *d = accumulate ? *d + acc : acc;
  }
}

// wT[(co * 9 + 8 - tap)][ci] <- w[(ci * 9 + tap)][co]
__global__ void transpose_packed_kernel(const float* __restrict__ wp, float* __restrict__ wt, int Cout, int Cin, int CoutP,
                                        int CinP) {
  const int total = Cout * 9 * CinP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % CinP, t = (i / CinP) % 9, co = i / (CinP * 9);
    wt[i] = ci < Cin ? wp[(size_t)(ci * 9 + (8 - t)) * CoutP + co] : 0.f;
  }
}
}  // namespace
}  // namespace b2f

extern "C" int b2f_conv3x3_transpose_packed(const float* w_packed, float* wt_packed, int Cout, int Cin, b2f_stream_t stream) {
  if (!w_packed || !wt_packed || Cout <= 0 || Cin <= 0) return fail(B2F_EINVAL, "conv3x3_transpose_packed: bad argument");
  const int CoutP = (Cout + 63) / 64 * 64, CinP = (Cin + 63) / 64 * 64;
  transpose_packed_kernel<<<ew_grid((int64_t)Cout * 9 * CinP, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w_packed, wt_packed, Cout, Cin, CoutP, CinP);
  B2F_CHECK_LAUNCH("transpose_packed_kernel");
  return B2F_OK;
}

extern "C" int b2f_conv3x3_backward_data(const float* gout, int64_t gout_batch_stride, const float* wt_packed,
                                         const float* act, int64_t act_batch_stride, float* gin, int64_t gin_batch_stride,
                                         int accumulate, int B, int Cin, int H, int W, int Cout, int stride,
                                         float leaky_slope, b2f_stream_t stream) {
  if (!gout || !wt_packed || !gin) return fail(B2F_EINVAL, "conv3x3_backward_data: NULL gout / wt_packed / gin");
  if (B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "conv3x3_backward_data: bad size");
  if (stride != 1 && stride != 2) return fail(B2F_EUNSUPPORTED, "conv3x3_backward_data: stride %d", stride);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const int64_t gbs = gout_batch_stride ? gout_batch_stride : (int64_t)Cout * Ho * Wo;
  const int64_t ibs = gin_batch_stride ? gin_batch_stride : (int64_t)Cin * H * W;
  const int64_t abs_ = act_batch_stride ? act_batch_stride : (int64_t)Cin * H * W;
  if (gbs < (int64_t)Cout * Ho * Wo || ibs < (int64_t)Cin * H * W || (act && abs_ < (int64_t)Cin * H * W))
    return fail(B2F_EINVAL, "conv3x3_backward_data: batch stride smaller than one item");
  if (!aligned4(gout) || !aligned4(gin) || !aligned4(wt_packed) || (act && !aligned4(act)))
    return fail(B2F_EALIGN, "conv3x3_backward_data: misaligned pointer");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int CinP = (Cin + 63) / 64 * 64;
  const bool tma_ok = stride == 1 && (W % 4) == 0 && aligned16(gout) && gbs % 4 == 0 && aligned16(wt_packed) &&
                      aligned16(gin) && ibs % 4 == 0 && get_encode_fn() != nullptr;
  if (tma_ok) {
    // the forward kernel with the roles of the channel counts exchanged: "input" = gout (Cout planes), "output" = gin
    cv3::Args a{nullptr, gin, nullptr, ibs, 0, Cout, Cin, H, W, 1, leaky_slope, act, abs_, accumulate};
    if (!act) a.slope = 1.f;
    const bool n64 = Cin % 64 == 0 || Cin > 96;
    if (Cin <= 16) return launch_conv<2, 1, 4>(gout, gbs, wt_packed, CinP, a, B, Cout, H, W, st);
    return n64 ? launch_conv<8, 1, 8>(gout, gbs, wt_packed, CinP, a, B, Cout, H, W, st)
               : launch_conv<4, 1, 8>(gout, gbs, wt_packed, CinP, a, B, Cout, H, W, st);
  }
  conv3x3_dgrad_generic<<<ew_grid((int64_t)B * Cin * H * W, 128), 128, 0, st>>>(gout, gbs, wt_packed, act, abs_, gin, ibs,
                                                                               accumulate, B, Cout, Ho, Wo, Cin, CinP, H, W,
                                                                               stride, leaky_slope);
  B2F_CHECK_LAUNCH("conv3x3_dgrad_generic");
  return B2F_OK;
}

extern "C" int b2f_avgpool2x2_forward(const float* x, float* out, int B, int C, int H, int W, b2f_stream_t stream) {
  if (!x || !out || B < 0 || C <= 0 || H < 2 || W < 2) return fail(B2F_EINVAL, "avgpool2x2_forward: bad argument");
  if ((W & 1) || (reinterpret_cast<uintptr_t>(x) & 7u)) return fail(B2F_EUNSUPPORTED, "avgpool2x2_forward: odd width or misaligned input");
  if (B == 0) return B2F_OK;
  const int Ho = H / 2, Wo = W / 2;
  avgpool2_kernel<<<ew_grid((int64_t)B * C * Ho * Wo, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, out, (int64_t)B * C, H, W, Ho, Wo);
  B2F_CHECK_LAUNCH("avgpool2_kernel");
  return B2F_OK;
}

extern "C" int b2f_upsample_bilinear2x_forward(const float* x, int64_t x_batch_stride, int B, int C, int H, int W,
                                               float* const* outs, const int64_t* out_batch_strides, int n_outs,
                                               float mul, b2f_stream_t stream) {
  if (!x || !outs || n_outs < 1 || n_outs > 3 || B < 0 || C <= 0 || H <= 0 || W <= 0)
    return fail(B2F_EINVAL, "upsample_bilinear2x_forward: bad argument");
  if (B == 0) return B2F_OK;
  UpDst d{};
  for (int i = 0; i < n_outs; ++i) {
    if (!outs[i]) return fail(B2F_EINVAL, "upsample_bilinear2x_forward: outs[%d] is NULL", i);
    d.p[i] = outs[i];
    d.bs[i] = (out_batch_strides && out_batch_strides[i]) ? out_batch_strides[i] : (int64_t)C * 4 * H * W;
  }
  const int64_t xbs = x_batch_stride ? x_batch_stride : (int64_t)C * H * W;
  upsample_bilinear2_kernel<<<ew_grid((int64_t)B * C * 4 * H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, xbs, d, B, C, H, W, mul);
  B2F_CHECK_LAUNCH("upsample_bilinear2_kernel");
  return B2F_OK;
}

extern "C" int b2f_upsample_nearest_forward(const float* x, float* out, int B, int C, int H, int W, int scale,
                                            b2f_stream_t stream) {
  if (!x || !out || B < 0 || C <= 0 || H <= 0 || W <= 0 || scale < 1) return fail(B2F_EINVAL, "upsample_nearest_forward: bad argument");
  if (B == 0) return B2F_OK;
  upsample_nearest_kernel<<<ew_grid((int64_t)B * C * H * W * scale * scale, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, out, (int64_t)B * C, H, W, scale);
  B2F_CHECK_LAUNCH("upsample_nearest_kernel");
  return B2F_OK;
}

extern "C" int b2f_softmax_channels_forward(const float* x, float* out, int B, int C, int H, int W, b2f_stream_t stream) {
  if (!x || !out || B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(B2F_EINVAL, "softmax_channels_forward: bad argument");
  if (B == 0) return B2F_OK;
  softmax_channels_kernel<<<ew_grid((int64_t)B * H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, out, B, C, (int64_t)H * W);
  B2F_CHECK_LAUNCH("softmax_channels_kernel");
  return B2F_OK;
}
