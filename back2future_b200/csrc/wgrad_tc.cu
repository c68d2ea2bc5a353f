// SpatialConvolution:accGradParameters (3 x 3, stride 1, pad 1) on the 5th-generation tensor cores for sm_100a.
//
//     gw[ci, tap, co] += sum over pixels p   X[p + tap, ci] * G[p, co]
//
// is a GEMM whose CONTRACTION index is the pixel.  The tensor-core forward and input-gradient kernels (conv_tc.cu)
// already keep activations and gradients channel-minor, (B, H, W, Cp) as (hi, lo) pairs: one row = the 32 channels of
// one pixel.  Read as MN-MAJOR operands (rows = K, tools/ubench/tc_gemm_mn.cu) those tensors ARE the operands:
//     A = the activation patch of a pixel tile, rows = pixels (K), columns = input channels (M = 128),
//     B = the gradient tile,                    rows = pixels (K), columns = output channels (N = Cout),
// and the tap (ky, kx) is, once more, a start offset of the A descriptor -- kx rows inside the patch.  tcgen05 accepts
// MN-major tf32 only in the SWIZZLE_128B_BASE32B shared-memory layout (32-byte chunks ^ (row & 3)), which TMA writes
// with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; LBO = byte distance of the 32-channel blocks, SBO = 512 (four pixel rows).
// Three kind::tf32 passes of the (hi, lo) split: X_hi G_hi + X_lo G_hi + X_hi G_lo, fp32 accumulation in TMEM.
//
// CTA = (128-input-channel chunk, kernel row ky, a share of the pixel tiles).  Its accumulators are the three taps
// (ky, 0..2) x N columns of TMEM (384 of 512 at N = 128) and stay there over ALL its tiles; one epilogue at the end adds
// them to the packed weight gradient with red.global.add (as the FFMA kernel does with its partial sums).  A pixel tile
// is 2 rows x 16 pixels: the patch of kernel row ky is the tile's two image rows shifted by ky - 1, 18 pixels wide
// (x0 - 1 .. x0 + 16; out-of-image pixels read as zero = the convolution's padding), the gradient tile is 2 x 16; per
// image row the two K = 8 steps of the patch start at rows 18 yl + kx (+ 8), those of the gradient tile at 16 yl (+ 8).
// Warp 0 lane 0 issues the TMA loads (3-stage ring, up to 2 x (4 + 4) boxes per stage), warp 1 lane 0 the MMAs (36 per
// stage), then all four warps read TMEM.
#include "tma.cuh"

#include <algorithm>

namespace b2f {
namespace {
namespace wtc {

constexpr int TH = 2, TW = 16, PW = TW + 2;
constexpr int XBLK = 5120;                  // one 32-channel block of the patch: 36 rows x 128 B, padded to a multiple of 1024
constexpr int GBLK = TH * TW * 128;         // one 32-channel block of the gradient tile: 4096
constexpr int MB = 4;                       // 32-channel blocks of M (128 input channels) and, at most, of N
constexpr int STAGE = 2 * MB * XBLK + 2 * MB * GBLK;   // hi + lo of both operands: 73 728 B
constexpr int STAGES = 3;
constexpr int SMEM_BYTES = STAGES * STAGE + 256 + 1024;
constexpr int THREADS = 128;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// MN-major, SWIZZLE_128B_BASE32B (layout type 1), version 1
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Args {
  float* gw;          // packed [Cin * 9][CoutP]
  int Cin, Cout, CoutP, H, W, B;
  int nxb_total;      // 32-channel blocks of the activation tensor (CinP / 32)
  int ngb;            // 32-channel blocks of the gradient (Cout / 32)
  int tiles_x, tiles_y, ntiles, nsplit;
  int co0;            // first output channel of this launch's slice (Cout > 128 runs as slices)
};

__global__ void __launch_bounds__(THREADS, 1)
conv3x3_wgrad_tc(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                 const __grid_constant__ CUtensorMap tm_gh, const __grid_constant__ CUtensorMap tm_gl, const Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* done = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x / 3, ky = blockIdx.x % 3;
  const int xb0 = chunk * MB;                                   // first 32-channel block of this CTA's input channels
  const int nxb = min(MB, a.nxb_total - xb0);
  const int N = a.ngb * 32;
  const uint32_t tmem_cols = 3 * N <= 128 ? 128u : (3 * N <= 256 ? 256u : 512u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  // blocks of M beyond the tensor's channels are never loaded: give the MMA zeros to read there (their TMEM lanes
  // are not stored, but NaN payloads of uninitialised shared memory need not travel through the tensor core)
  if (nxb < MB) {
    for (int s = 0; s < STAGES; ++s)
      for (int p = 0; p < 2; ++p) {
        float4* z = reinterpret_cast<float4*>(smem + s * STAGE + p * MB * XBLK + nxb * XBLK);
        for (int i = threadIdx.x; i < (MB - nxb) * XBLK / 16; i += THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int my_tiles = (a.ntiles - (int)blockIdx.y + a.nsplit - 1) / a.nsplit;   // tiles blockIdx.y, + nsplit, ...

  if (warp == 0 && lane == 0) {
    // ---- TMA producer ----
    const uint32_t bytes = (uint32_t)(2 * nxb * (TH * PW * 128) + 2 * a.ngb * GBLK);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
      const int t = blockIdx.y + it * a.nsplit;
      const int tx = t % a.tiles_x, tq = t / a.tiles_x;
      const int ty = tq % a.tiles_y, b = tq / a.tiles_y;
      const int x0 = tx * TW, y0 = ty * TH;
      uint8_t* st = smem + s * STAGE;
      mbar_arrive_expect_tx(&full[s], bytes);
      for (int c = 0; c < nxb; ++c) {
        tma_load_4d(st + c * XBLK, &tm_xh, (xb0 + c) * 32, x0 - 1, y0 + ky - 1, b, &full[s]);
        tma_load_4d(st + MB * XBLK + c * XBLK, &tm_xl, (xb0 + c) * 32, x0 - 1, y0 + ky - 1, b, &full[s]);
      }
      uint8_t* gs = st + 2 * MB * XBLK;
      for (int c = 0; c < a.ngb; ++c) {
        tma_load_4d(gs + c * GBLK, &tm_gh, a.co0 + c * 32, x0, y0, b, &full[s]);
        tma_load_4d(gs + MB * GBLK + c * GBLK, &tm_gl, a.co0 + c * 32, x0, y0, b, &full[s]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer ----
    // instruction descriptor: D fp32, A and B TF32, both MN-major (bits 15, 16), N >> 3 at bit 17, M = 128 (>> 4 at bit 24)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    uint32_t started = 0;                                      // bit kx: the tap's columns hold a partial sum
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t xh = smem_u32(smem + s * STAGE), xl = xh + MB * XBLK;
      const uint32_t gh = xh + 2 * MB * XBLK, gl = gh + MB * GBLK;
#pragma unroll 1
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t dcol = tmem + (uint32_t)(kx * N);
#pragma unroll
        for (int p = 0; p < 3; ++p) {                          // hi * hi, lo * hi, hi * lo
          const uint32_t pa = (p == 1 ? xl : xh), pb = (p == 2 ? gl : gh);
#pragma unroll
          for (int yl = 0; yl < TH; ++yl)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              mma_tf32(dcol, desc_mn(pa + (uint32_t)(yl * PW + kx + 8 * h) * 128u, XBLK),
                       desc_mn(pb + (uint32_t)(yl * TW + 8 * h) * 128u, GBLK), idesc, (started >> kx) & 1u);
              started |= 1u << kx;
            }
        }
      }
      umma_commit(&empty[s]);
    }
    umma_commit(done);
  }
  __syncwarp();

  // ---- epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 = input channels of the chunk ----
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int ci = chunk * 128 + 32 * warp + lane;
  if (my_tiles > 0) {
    for (int kx = 0; kx < 3; ++kx) {
      float* dst = a.gw + ((size_t)ci * 9 + (ky * 3 + kx)) * a.CoutP + a.co0;
      for (int n0 = 0; n0 < N; n0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(kx * N + n0)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ci < a.Cin) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (a.co0 + n0 + j < a.Cout) atomicAdd(dst + n0 + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

// gb[co] += sum over pixels of the PLANAR output gradient (one block per channel and batch item)
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ g, int64_t gbs, float* __restrict__ gb, int64_t hw) {
  const int co = blockIdx.x, b = blockIdx.y;
  const float* p = g + (size_t)b * gbs + (size_t)co * hw;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < hw; i += 256) s += __ldg(p + i);
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(gb + co, t);
  }
}

// The same sum from the channel-minor (hi, lo) output gradient, (pixels, Cp): a warp reads one 128-byte line of 32
// channels per pixel, eight pixels per block and trip
__global__ void __launch_bounds__(256) bias_grad_nhwc_kernel(const float* __restrict__ gh, const float* __restrict__ gl,
                                                            float* __restrict__ gb, int64_t npix, int Cp, int Cout) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = 32 * blockIdx.x + lane;
  float s = 0.f;
  for (int64_t p = (int64_t)blockIdx.y * 8 + w; p < npix; p += (int64_t)gridDim.y * 8) {
    const size_t o = (size_t)p * Cp + c;
    s += __ldg(gh + o) + __ldg(gl + o);
  }
  __shared__ float part[8][32];
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < Cout) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][lane];
    atomicAdd(gb + c, t);
  }
}

int make_tmap4_atom32(CUtensorMap* tm, const float* base, const uint64_t dims[4], const uint64_t strides_elems[3], const uint32_t box[4]) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(B2F_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_elems[0] * 4, strides_elems[1] * 4, strides_elems[2] * 4};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2F_EINVAL, "cuTensorMapEncodeTiled (128B_ATOM_32B) failed with CUresult %d", (int)r);
  return B2F_OK;
}

}  // namespace wtc
}  // namespace
}  // namespace b2f

using namespace b2f;

extern "C" int b2f_conv3x3_tc_backward_weights(const float* x_hi, const float* x_lo, int Cx, const float* g_hi, const float* g_lo,
                                               const float* g_planar, int64_t g_planar_batch_stride, float* gw_packed,
                                               float* gbias, int B, int Cin, int H, int W, int Cout, b2f_stream_t stream) {
  if (!x_hi || !x_lo || !g_hi || !g_lo || !gw_packed) return fail(B2F_EINVAL, "conv3x3_tc_backward_weights: NULL operand");
  if (B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || Cx < Cin) return fail(B2F_EINVAL, "conv3x3_tc_backward_weights: bad size");
  if (!aligned16(x_hi) || !aligned16(x_lo) || !aligned16(g_hi) || !aligned16(g_lo)) return fail(B2F_EALIGN, "conv3x3_tc_backward_weights: operands must be 16-byte aligned");
  if (get_encode_fn() == nullptr) return fail(B2F_EUNSUPPORTED, "conv3x3_tc_backward_weights: cuTensorMapEncodeTiled not available");
  if (B == 0) return B2F_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int CxP = (Cx + 31) / 32 * 32, CoutP32 = (Cout + 31) / 32 * 32;
  CUtensorMap txh, txl, tgh, tgl;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)CxP, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)CxP, (uint64_t)CxP * W, (uint64_t)CxP * W * H};
    const uint32_t box[4] = {32, (uint32_t)wtc::PW, (uint32_t)wtc::TH, 1};
    if ((rc = wtc::make_tmap4_atom32(&txh, x_hi, dims, str, box))) return rc;
    if ((rc = wtc::make_tmap4_atom32(&txl, x_lo, dims, str, box))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)CoutP32, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)CoutP32, (uint64_t)CoutP32 * W, (uint64_t)CoutP32 * W * H};
    const uint32_t box[4] = {32, (uint32_t)wtc::TW, (uint32_t)wtc::TH, 1};
    if ((rc = wtc::make_tmap4_atom32(&tgh, g_hi, dims, str, box))) return rc;
    if ((rc = wtc::make_tmap4_atom32(&tgl, g_lo, dims, str, box))) return rc;
  }
  static thread_local int attr_dev = -1;
  int dev = 0;
  B2F_CUDA_TRY(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    B2F_CUDA_TRY(cudaFuncSetAttribute(wtc::conv3x3_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::SMEM_BYTES));
    attr_dev = dev;
  }
  // output-channel slices of <= 128 (the accumulators of three taps are 3 N TMEM columns)
  for (int co0 = 0; co0 < Cout; co0 += 128) {
    wtc::Args a{};
    a.gw = gw_packed;
    a.Cin = Cin; a.Cout = Cout; a.CoutP = (Cout + 63) / 64 * 64;
    a.H = H; a.W = W; a.B = B;
    a.nxb_total = (Cin + 31) / 32;          // blocks that hold channels of THIS convolution (the tensor may be wider)
    a.ngb = (std::min(128, Cout - co0) + 31) / 32;
    a.co0 = co0;
    a.tiles_x = (W + wtc::TW - 1) / wtc::TW;
    a.tiles_y = (H + wtc::TH - 1) / wtc::TH;
    a.ntiles = B * a.tiles_x * a.tiles_y;
    const int nchunk = (a.nxb_total + wtc::MB - 1) / wtc::MB;
    // one CTA per SM (it owns up to all 512 TMEM columns); at least 8 tiles per CTA so the 3-stage ring has something to overlap
    a.nsplit = std::max(1, std::min(num_sms() / (3 * nchunk), std::max(1, a.ntiles / 8)));
    dim3 grid(3 * nchunk, a.nsplit);
    wtc::conv3x3_wgrad_tc<<<grid, wtc::THREADS, wtc::SMEM_BYTES, st>>>(txh, txl, tgh, tgl, a);
    B2F_CHECK_LAUNCH("conv3x3_wgrad_tc");
  }
  if (gbias && !g_planar) {
    const int64_t npix = (int64_t)B * H * W;
    const int ny = (int)std::max<int64_t>(1, std::min<int64_t>((npix + 63) / 64, 4 * num_sms()));
    wtc::bias_grad_nhwc_kernel<<<dim3(CoutP32 / 32, ny), 256, 0, st>>>(g_hi, g_lo, gbias, npix, CoutP32, Cout);
    B2F_CHECK_LAUNCH("bias_grad_nhwc_kernel");
  } else if (gbias) {
    const int64_t gbs = g_planar_batch_stride ? g_planar_batch_stride : (int64_t)Cout * H * W;
    wtc::bias_grad_kernel<<<dim3(Cout, B), 256, 0, st>>>(g_planar, gbs, gbias, (int64_t)H * W);
    B2F_CHECK_LAUNCH("bias_grad_kernel");
  }
  return B2F_OK;
}
