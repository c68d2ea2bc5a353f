// Fused criterion kernels for sm_100a: one pass per criterion call produces the loss scalar and every
// gradient the reference returns.  All of them are pure streaming stencils (HBM-bound); the reference
// runs 10-100 generic THC kernels plus host-built coordinate grids per call.
//
// Loss reduction is deterministic: per-thread float partial -> per-block double partial -> the last
// block to finish (ticket counter) sums the block partials in index order and applies the normaliser.
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace b2f {
namespace {

constexpr int kThreads = 128;   // one thread per pixel of a row segment; grid = (x tiles, row, batch)

struct LossOut {
  double* partials;   // [gridDim.x]
  unsigned* counter;  // zeroed before launch
  double* result;     // scratch result (device)
  double* loss_dev;   // optional user device pointer
  double scale;       // normaliser applied once at the end
};

__device__ __forceinline__ void finish_loss(float local, const LossOut& lo) {
  __shared__ double s_warp[kThreads / 32];
  __shared__ bool s_last;
  double v = (double)local;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) t += s_warp[i];
    lo.partials[bid] = t;
    __threadfence();
    const unsigned ticket = atomicAdd(lo.counter, 1u);
    s_last = (ticket == nblocks - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last block: fixed-order sum of the block partials
  double t = 0.0;
  for (unsigned i = threadIdx.x; i < nblocks; i += kThreads) t += __ldcg(lo.partials + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  if (lane == 0) s_warp[warp] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) tot += s_warp[i];
    tot *= lo.scale;
    *lo.result = tot;
    if (lo.loss_dev) *lo.loss_dev = tot;
    *lo.counter = 0u;   // self-cleaning ticket: the next call on this stream needs no memset
  }
}

// Host side of the loss delivery.  The scratch for the block partials is owned by the library: one
// buffer per (host thread, device, stream), grown on demand and reused by later calls on that stream
// (calls on one stream are ordered, so reuse is safe); b2f_release_scratch() frees the calling thread's
// buffers.  (A per-call cudaMallocAsync/cudaFreeAsync pair cost milliseconds: the default pool's release
// threshold is 0, so every synchronisation handed the memory back to the driver.)
//
// Lifetime rules that make the scratch safe under CUDA graphs (a captured kernel node has its partials / ticket /
// result pointers baked in):
//   * nothing is ever freed before b2f_release_scratch(): a buffer that has to grow is RETIRED, not freed, and no
//     entry is evicted, so a graph captured earlier never replays into freed memory;
//   * a call issued while its stream is CAPTURING gets a slice of its own from the graph arena reserved with
//     b2f_reserve_scratch() (bump-allocated, never reused), so a replay on any stream cannot race with eager
//     criterion calls -- or with another captured call -- on the ticket counter.  Without a reserved arena the
//     captured call falls back to the capture stream's buffer (which must then already be large enough:
//     allocating is illegal during capture) and the caller must not run eager criterion calls on that stream
//     while the graph replays.
struct ScratchEntry {
  int dev;
  cudaStream_t st;
  void* mem;
  size_t bytes;
};
struct ScratchCache {
  std::vector<ScratchEntry> e;
  std::vector<void*> retired;          // outgrown buffers, kept alive for graphs captured earlier
  struct Arena { int dev; char* mem; size_t bytes, used; };
  std::vector<Arena> arenas;           // graph arenas (b2f_reserve_scratch), one or more per device
  ~ScratchCache() {}   // device memory is released explicitly (b2f_release_scratch) or at process exit
};
thread_local ScratchCache g_scratch;

int get_scratch(size_t bytes, cudaStream_t st, void** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  ScratchCache& c = g_scratch;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) cap = cudaStreamCaptureStatusNone;
  if (cap == cudaStreamCaptureStatusActive) {
    const size_t need = (bytes + 255) & ~(size_t)255;
    for (auto& a : c.arenas)
      if (a.dev == dev && a.bytes - a.used >= need) {
        *out = a.mem + a.used;      // zeroed at reserve time; every kernel leaves its ticket at zero
        a.used += need;
        return B2F_OK;
      }
  }
  int slot = -1;
  for (size_t i = 0; i < c.e.size(); ++i)
    if (c.e[i].dev == dev && c.e[i].st == st) slot = (int)i;
  if (slot < 0) {
    slot = (int)c.e.size();
    c.e.push_back({dev, st, nullptr, 0});
  }
  ScratchEntry& s = c.e[slot];
  if (s.bytes < bytes) {
    if (cap == cudaStreamCaptureStatusActive)
      return fail(B2F_ENOMEM, "criterion scratch: %zu bytes needed while the stream is capturing; call "
                  "b2f_reserve_scratch() before the capture (allocation is illegal during capture)", bytes);
    if (s.mem) {
      c.retired.push_back(s.mem);   // a graph captured earlier may still point into it
      s.mem = nullptr;
      s.bytes = 0;
    }
    size_t want = bytes < (size_t)(1 << 20) ? (size_t)(1 << 20) : bytes * 2;
    e = cudaMalloc(&s.mem, want);
    if (e != cudaSuccess) {
      s.mem = nullptr;
      return fail(B2F_ENOMEM, "criterion scratch: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    s.bytes = want;
    // the ticket counter starts at zero and every kernel leaves it at zero (finish_loss)
    e = cudaMemsetAsync(s.mem, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(loss scratch)");
  }
  *out = s.mem;
  return B2F_OK;
}

struct LossScratch {
  LossOut lo{};
  cudaStream_t st = nullptr;

  int begin(int blocks, double scale, double* loss_dev, cudaStream_t stream) {
    st = stream;
    const size_t bytes = (size_t)blocks * sizeof(double) + 2 * sizeof(double);
    void* mem = nullptr;
    int rc = get_scratch(bytes, st, &mem);
    if (rc) return rc;
    lo.result = reinterpret_cast<double*>(mem);
    lo.counter = reinterpret_cast<unsigned*>(lo.result + 1);
    lo.partials = lo.result + 2;
    lo.loss_dev = loss_dev;
    lo.scale = scale;
    return B2F_OK;
  }
  int end(double* loss_host) {
    if (loss_host) {
      cudaError_t e = cudaMemcpyAsync(loss_host, lo.result, sizeof(double), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return cuda_fail(e, "loss read-back");
    }
    return B2F_OK;
  }
};

}  // namespace

int release_scratch_for_thread() {
  ScratchCache& c = g_scratch;
  for (auto& en : c.e)
    if (en.mem) cudaFree(en.mem);
  for (void* p : c.retired) cudaFree(p);
  for (auto& a : c.arenas) cudaFree(a.mem);
  c.e.clear();
  c.retired.clear();
  c.arenas.clear();
  return B2F_OK;
}

int reserve_scratch_for_thread(size_t bytes) {
  if (bytes == 0) return fail(B2F_EINVAL, "b2f_reserve_scratch: bytes must be positive");
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  void* mem = nullptr;
  e = cudaMalloc(&mem, bytes);
  if (e != cudaSuccess) return fail(B2F_ENOMEM, "b2f_reserve_scratch: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(mem, 0, bytes);
  if (e != cudaSuccess) { cudaFree(mem); return cuda_fail(e, "cudaMemset(graph arena)"); }
  g_scratch.arenas.push_back({dev, reinterpret_cast<char*>(mem), bytes, 0});
  return B2F_OK;
}

namespace {

// =======================================================================================
// OBCC / OBGCC  (criterions/OBCCriterion.lua, criterions/OBGCCriterion.lua), F = 3
// =======================================================================================
struct ObArgs {
  const float* flow;
  const float* bflow;   // flow used for the past frame's mask (== flow unless past_flow)
  const float* occ;
  const float* warp[2];  // [0] past frame (f=1), [1] future frame (f=2)
  const float* target;
  float* g_occ;
  float* g_warp[2];
  int B, C, h, w;
  float eps2, penalty_out, alpha, beta, gamma, scale, norm;
  int grad_check;
};

// One pixel, any channel count: the reference's formulas channel by channel (loads, arithmetic and the gradient
// store of a channel follow each other, so the memory round trips of the channels are serialised).
template <int PEN, bool GT>
__device__ __forceinline__ void ob_pixel_generic(const ObArgs& a, int b, int y, int x, int64_t hw, float& loss) {
    const int64_t o = (int64_t)y * a.w + x;
    const int w = a.w, h = a.h;
#pragma unroll
    for (int fr = 0; fr < 2; ++fr) {
      // fr 0: past frame, k = -1, occ channel 2 (index 1); fr 1: future, k = +1, occ channel 1
      const float k = fr == 0 ? -1.f : 1.f;
      const int oc = fr == 0 ? 1 : 0;
      const float* fl = fr == 0 ? a.bflow : a.flow;
      const float occv = __ldg(a.occ + ((int64_t)b * 2 + oc) * hw + o);
      bool m = true;
      if (!a.grad_check) {
        // tcoord = fl(coord + fl(fl(k*flow)*scale)), 1-based coords (OBCCriterion.lua:81-100, Q14)
        const float fx = __ldg(fl + ((int64_t)b * 2) * hw + o);
        const float fy = __ldg(fl + ((int64_t)b * 2 + 1) * hw + o);
        const float tx = __fadd_rn((float)(x + 1), __fmul_rn(__fmul_rn(fx, k), a.scale));
        const float ty = __fadd_rn((float)(y + 1), __fmul_rn(__fmul_rn(fy, k), a.scale));
        m = (tx >= 1.f) && (ty >= 1.f) && (tx <= (float)w) && (ty <= (float)h);
      }
      const float* img = a.warp[fr];
      float* gw = a.g_warp[fr];
      float e = 0.f, ex = 0.f, ey = 0.f, exm = 0.f, eym = 0.f;
      const float gscale = m ? occv * a.norm : 0.f;
      for (int c = 0; c < a.C; ++c) {
        const int64_t p = ((int64_t)b * a.C + c) * hw + o;
        const float iv = __ldg(img + p), tv = __ldg(a.target + p);
        const float d = iv - tv;
        e += pen_apply<PEN>(d, a.eps2);
        float gi;
        if (GT) {
          const float dgx = (x < w - 1) ? (__ldg(img + p + 1) - iv) - (__ldg(a.target + p + 1) - tv) : 0.f;
          const float dgy = (y < h - 1) ? (__ldg(img + p + w) - iv) - (__ldg(a.target + p + w) - tv) : 0.f;
          ex += pen_apply<PEN>(dgx, a.eps2);
          ey += pen_apply<PEN>(dgy, a.eps2);
          gi = pen_der<PEN>(d, a.eps2) * a.alpha;
          gi -= pen_der<PEN>(dgy, a.eps2) * a.gamma;
          if (y > 0) {
            const float dm = (iv - __ldg(img + p - w)) - (tv - __ldg(a.target + p - w));
            gi += pen_der<PEN>(dm, a.eps2) * a.gamma;
            eym += pen_apply<PEN>(dm, a.eps2);
          }
          gi -= pen_der<PEN>(dgx, a.eps2) * a.beta;
          if (x > 0) {
            const float dm = (iv - __ldg(img + p - 1)) - (tv - __ldg(a.target + p - 1));
            gi += pen_der<PEN>(dm, a.eps2) * a.beta;
            exm += pen_apply<PEN>(dm, a.eps2);
          }
        } else {
          gi = pen_der<PEN>(d, a.eps2);
        }
        if (gw) gw[p] = gi * gscale;
      }
      // forward energy (alpha is not applied here: OBGCCriterion.lua:97)
      float tmp = GT ? e + ex * a.beta + ey * a.gamma : e;
      tmp *= occv;
      loss += m ? tmp : a.penalty_out;
      // occlusion gradient (Q6, Q7)
      if (a.g_occ) {
        float buf = e;
        if (GT) {
          buf = e * a.alpha;
          buf -= ey * a.gamma;
          buf += eym * a.gamma;
          buf -= ex * a.beta;
          buf += exm * a.beta;
        }
        buf = m ? buf : a.penalty_out;
        a.g_occ[((int64_t)b * 2 + oc) * hw + o] = buf * a.norm;
      }
    }
}

// One pixel, C = 3 (every criterion call of the model): ALL loads of both frames are issued first -- 2 occlusion,
// 4 flow, 3 target and 6 warped values, plus their four neighbours with the gradient terms -- so a pixel costs
// one memory round trip instead of one per channel and frame; the arithmetic and its order are unchanged.
template <int PEN, bool GT>
__device__ __forceinline__ void ob_pixel_c3(const ObArgs& a, int b, int y, int x, int64_t hw, float& loss) {
  const int64_t o = (int64_t)y * a.w + x;
  const int w = a.w, h = a.h;
  float occv[2], fx[2] = {0.f, 0.f}, fy[2] = {0.f, 0.f};
#pragma unroll
  for (int fr = 0; fr < 2; ++fr) {
    const int oc = fr == 0 ? 1 : 0;
    const float* fl = fr == 0 ? a.bflow : a.flow;
    occv[fr] = __ldg(a.occ + ((int64_t)b * 2 + oc) * hw + o);
    if (!a.grad_check) {
      fx[fr] = __ldg(fl + ((int64_t)b * 2) * hw + o);
      fy[fr] = __ldg(fl + ((int64_t)b * 2 + 1) * hw + o);
    }
  }
  // centre and (GT) the four neighbours of target and of both warped frames; index 0 = target, 1/2 = frames
  float v0[3][3], vxp[3][3], vyp[3][3], vxm[3][3], vym[3][3];
#pragma unroll
  for (int s3 = 0; s3 < 3; ++s3) {
    const float* src = s3 == 0 ? a.target : a.warp[s3 - 1];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* q = src + ((int64_t)b * 3 + c) * hw + o;
      v0[s3][c] = __ldg(q);
      if (GT) {
        vxp[s3][c] = (x < w - 1) ? __ldg(q + 1) : 0.f;
        vyp[s3][c] = (y < h - 1) ? __ldg(q + w) : 0.f;
        vxm[s3][c] = (x > 0) ? __ldg(q - 1) : 0.f;
        vym[s3][c] = (y > 0) ? __ldg(q - w) : 0.f;
      }
    }
  }
  float gout[2][3], gocc[2];
#pragma unroll
  for (int fr = 0; fr < 2; ++fr) {
    // fr 0: past frame, k = -1, occ channel 2 (index 1); fr 1: future, k = +1, occ channel 1
    const float k = fr == 0 ? -1.f : 1.f;
    bool m = true;
    if (!a.grad_check) {
      // tcoord = fl(coord + fl(fl(k*flow)*scale)), 1-based coords (OBCCriterion.lua:81-100, Q14)
      const float tx = __fadd_rn((float)(x + 1), __fmul_rn(__fmul_rn(fx[fr], k), a.scale));
      const float ty = __fadd_rn((float)(y + 1), __fmul_rn(__fmul_rn(fy[fr], k), a.scale));
      m = (tx >= 1.f) && (ty >= 1.f) && (tx <= (float)w) && (ty <= (float)h);
    }
    float e = 0.f, ex = 0.f, ey = 0.f, exm = 0.f, eym = 0.f;
    const float gscale = m ? occv[fr] * a.norm : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float iv = v0[fr + 1][c], tv = v0[0][c];
      const float d = iv - tv;
      e += pen_apply<PEN>(d, a.eps2);
      float gi;
      if (GT) {
        const float dgx = (x < w - 1) ? (vxp[fr + 1][c] - iv) - (vxp[0][c] - tv) : 0.f;
        const float dgy = (y < h - 1) ? (vyp[fr + 1][c] - iv) - (vyp[0][c] - tv) : 0.f;
        ex += pen_apply<PEN>(dgx, a.eps2);
        ey += pen_apply<PEN>(dgy, a.eps2);
        gi = pen_der<PEN>(d, a.eps2) * a.alpha;
        gi -= pen_der<PEN>(dgy, a.eps2) * a.gamma;
        if (y > 0) {
          const float dm = (iv - vym[fr + 1][c]) - (tv - vym[0][c]);
          gi += pen_der<PEN>(dm, a.eps2) * a.gamma;
          eym += pen_apply<PEN>(dm, a.eps2);
        }
        gi -= pen_der<PEN>(dgx, a.eps2) * a.beta;
        if (x > 0) {
          const float dm = (iv - vxm[fr + 1][c]) - (tv - vxm[0][c]);
          gi += pen_der<PEN>(dm, a.eps2) * a.beta;
          exm += pen_apply<PEN>(dm, a.eps2);
        }
      } else {
        gi = pen_der<PEN>(d, a.eps2);
      }
      gout[fr][c] = gi * gscale;
    }
    // forward energy (alpha is not applied here: OBGCCriterion.lua:97)
    float tmp = GT ? e + ex * a.beta + ey * a.gamma : e;
    tmp *= occv[fr];
    loss += m ? tmp : a.penalty_out;
    // occlusion gradient (Q6, Q7)
    float buf = e;
    if (GT) {
      buf = e * a.alpha;
      buf -= ey * a.gamma;
      buf += eym * a.gamma;
      buf -= ex * a.beta;
      buf += exm * a.beta;
    }
    buf = m ? buf : a.penalty_out;
    gocc[fr] = buf * a.norm;
  }
#pragma unroll
  for (int fr = 0; fr < 2; ++fr) {
    if (a.g_warp[fr]) {
#pragma unroll
      for (int c = 0; c < 3; ++c) a.g_warp[fr][((int64_t)b * 3 + c) * hw + o] = gout[fr][c];
    }
    if (a.g_occ) a.g_occ[((int64_t)b * 2 + (fr == 0 ? 1 : 0)) * hw + o] = gocc[fr];
  }
}

// OBCC (no gradient terms), C = 3, software-pipelined over the block's rows: the 15 loads of the NEXT row are issued
// before the current row is evaluated (ncu had the one-row-at-a-time form at 53 % long-scoreboard stalls, 45 %
// occupancy: two rows of loads in flight per thread instead of one).  Same arithmetic and order as ob_pixel_c3.
struct ObRaw {
  float occv[2], fx[2], fy[2], v0[3][3];
};
__device__ __forceinline__ void ob_fetch_c3(const ObArgs& a, int b, int y, int x, int64_t hw, ObRaw& r) {
  const int64_t o = (int64_t)y * a.w + x;
#pragma unroll
  for (int fr = 0; fr < 2; ++fr) {
    const int oc = fr == 0 ? 1 : 0;
    const float* fl = fr == 0 ? a.bflow : a.flow;
    r.occv[fr] = __ldg(a.occ + ((int64_t)b * 2 + oc) * hw + o);
    r.fx[fr] = r.fy[fr] = 0.f;
    if (!a.grad_check) {
      r.fx[fr] = __ldg(fl + ((int64_t)b * 2) * hw + o);
      r.fy[fr] = __ldg(fl + ((int64_t)b * 2 + 1) * hw + o);
    }
  }
#pragma unroll
  for (int s3 = 0; s3 < 3; ++s3) {
    const float* src = s3 == 0 ? a.target : a.warp[s3 - 1];
#pragma unroll
    for (int c = 0; c < 3; ++c) r.v0[s3][c] = __ldg(src + ((int64_t)b * 3 + c) * hw + o);
  }
}
template <int PEN>
__device__ __forceinline__ void ob_eval_c3(const ObArgs& a, int b, int y, int x, int64_t hw, const ObRaw& r, float& loss) {
  const int64_t o = (int64_t)y * a.w + x;
  const int w = a.w, h = a.h;
  float gout[2][3], gocc[2];
#pragma unroll
  for (int fr = 0; fr < 2; ++fr) {
    const float k = fr == 0 ? -1.f : 1.f;
    bool m = true;
    if (!a.grad_check) {
      const float tx = __fadd_rn((float)(x + 1), __fmul_rn(__fmul_rn(r.fx[fr], k), a.scale));
      const float ty = __fadd_rn((float)(y + 1), __fmul_rn(__fmul_rn(r.fy[fr], k), a.scale));
      m = (tx >= 1.f) && (ty >= 1.f) && (tx <= (float)w) && (ty <= (float)h);
    }
    float e = 0.f;
    const float gscale = m ? r.occv[fr] * a.norm : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = r.v0[fr + 1][c] - r.v0[0][c];
      e += pen_apply<PEN>(d, a.eps2);
      gout[fr][c] = pen_der<PEN>(d, a.eps2) * gscale;
    }
    float tmp = e;
    tmp *= r.occv[fr];
    loss += m ? tmp : a.penalty_out;
    const float buf = m ? e : a.penalty_out;
    gocc[fr] = buf * a.norm;
  }
#pragma unroll
  for (int fr = 0; fr < 2; ++fr) {
    if (a.g_warp[fr]) {
#pragma unroll
      for (int c = 0; c < 3; ++c) a.g_warp[fr][((int64_t)b * 3 + c) * hw + o] = gout[fr][c];
    }
    if (a.g_occ) a.g_occ[((int64_t)b * 2 + (fr == 0 ? 1 : 0)) * hw + o] = gocc[fr];
  }
}

template <int PEN, bool GT>
__global__ void __launch_bounds__(kThreads)
ob_kernel(ObArgs a, LossOut lo) {
  const int64_t hw = (int64_t)a.h * a.w;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  const int b = blockIdx.z;
  float loss = 0.f;
  // a block walks rows blockIdx.y, blockIdx.y + gridDim.y, ...
  if (x < a.w) {
    if (a.C == 3 && !GT) {
      int y = blockIdx.y;
      ObRaw cur, nxt;
      if (y < a.h) ob_fetch_c3(a, b, y, x, hw, cur);
      for (; y < a.h; y += gridDim.y) {
        const int yn = y + (int)gridDim.y;
        if (yn < a.h) ob_fetch_c3(a, b, yn, x, hw, nxt);
        ob_eval_c3<PEN>(a, b, y, x, hw, cur, loss);
        cur = nxt;
      }
    } else if (a.C == 3) {
      for (int y = blockIdx.y; y < a.h; y += gridDim.y) ob_pixel_c3<PEN, GT>(a, b, y, x, hw, loss);
    } else {
      for (int y = blockIdx.y; y < a.h; y += gridDim.y) ob_pixel_generic<PEN, GT>(a, b, y, x, hw, loss);
    }
  }
  finish_loss(loss, lo);
}

// =======================================================================================
// SmoothnessCriterion / SecondOrderSmoothnessCriterion
// =======================================================================================
struct SmArgs {
  const float* in;
  const float* tgt;
  float* grad;
  int B, Cin, Ct, h, w;
  float eps2, cs, norm;
  int alias;        // order 1: reproduce the view-resize aliasing (only matters when Cin != Ct)
  int64_t n_dy;     // B*Ct*(h-1)*w, elements of the contiguous y-difference array
  int64_t n_dx;     // B*Ct*h*(w-1)
  int small;        // all flat indices fit in 32 bits
  // exact division of n < 2^31 by the loop-invariant h-1 / w-1: q = umulhi(n, mul) >> shr (d > 1), else n
  uint32_t mul_h1, shr_h1, mul_w1, shr_w1;
};

__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t d, uint32_t mul, uint32_t shr) {
  return d == 1u ? n : (__umulhi(n, mul) >> shr);
}

// order-1 edge weights -----------------------------------------------------------------
template <int CIN, int CT>
__device__ __forceinline__ float w1_y(const SmArgs& a, int b, int y, int x) {
  const int64_t hw = (int64_t)a.h * a.w;
  const int Cin = CIN ? CIN : a.Cin, Ct = CT ? CT : a.Ct;
  (void)Cin; (void)Ct;
  float s = 0.f;
  if (a.alias) {
    // igy(b,j,y,x) = flat(D_y)[((b*Cin+j)*h+y)*w+x], D_y contiguous (B,Ct,h-1,w)   (Q9)
    #pragma unroll
    for (int j = 0; j < Cin; ++j) {
      // idx = t*w + x with t = (b*Cin+j)*h + y, so idx % w == x and idx / w == t: only the small
      // row counter t has to be decomposed (32-bit)
      const unsigned t = ((unsigned)b * Cin + j) * a.h + y;
      float v = 0.f;
      if ((int64_t)t * a.w + x < a.n_dy) {
        const unsigned t2 = fast_div(t, (unsigned)(a.h - 1), a.mul_h1, a.shr_h1);   // = bb*Ct + cc
        const unsigned r = t - t2 * (unsigned)(a.h - 1);
        const float* p = a.tgt + (int64_t)t2 * hw + (int64_t)r * a.w + x;
        v = __ldg(p + a.w) - __ldg(p);
      }
      s += fabsf(v);
    }
    return __expf(-a.cs * (s / (float)Cin));
  }
  if (y < a.h - 1) {
#pragma unroll
    for (int c = 0; c < Ct; ++c) {
      const float* p = a.tgt + ((int64_t)b * Ct + c) * hw + (int64_t)y * a.w + x;
      s += fabsf(__ldg(p + a.w) - __ldg(p));
    }
  }
  return __expf(-a.cs * (s / (float)Ct));
}

template <int CIN, int CT>
__device__ __forceinline__ float w1_x(const SmArgs& a, int b, int y, int x) {
  const int64_t hw = (int64_t)a.h * a.w;
  const int Cin = CIN ? CIN : a.Cin, Ct = CT ? CT : a.Ct;
  (void)Cin; (void)Ct;
  float s = 0.f;
  if (a.alias) {
    #pragma unroll
    for (int j = 0; j < Cin; ++j) {
      const int64_t idx = (((int64_t)b * Cin + j) * a.h + y) * a.w + x;
      float v = 0.f;
      if (idx < a.n_dx) {
        // rows of the contiguous x-difference array are w-1 long: row = idx / (w-1), 32-bit when it fits
        unsigned rowi, xx;
        if (a.small) {
          rowi = fast_div((unsigned)idx, (unsigned)(a.w - 1), a.mul_w1, a.shr_w1);
          xx = (unsigned)idx - rowi * (unsigned)(a.w - 1);
        } else {
          rowi = (unsigned)(idx / (a.w - 1));
          xx = (unsigned)(idx - (int64_t)rowi * (a.w - 1));
        }
        // rowi = (bb*Ct + cc)*h + yy and the target rows are contiguous, so the source offset is rowi*w + xx
        const float* p = a.tgt + (int64_t)rowi * a.w + xx;
        v = __ldg(p + 1) - __ldg(p);
      }
      s += fabsf(v);
    }
    return __expf(-a.cs * (s / (float)Cin));
  }
  if (x < a.w - 1) {
#pragma unroll
    for (int c = 0; c < Ct; ++c) {
      const float* p = a.tgt + ((int64_t)b * Ct + c) * hw + (int64_t)y * a.w + x;
      s += fabsf(__ldg(p + 1) - __ldg(p));
    }
  }
  return __expf(-a.cs * (s / (float)Ct));
}

// CIN / CT: compile-time channel counts (2, 3 = every call of the model) or 0 = run-time.  With constant trip
// counts the loads of a pixel are straight-line code and the gradient stores are deferred to the end, so they
// all overlap (with run-time loops every channel costs its own memory round trip).
template <int PEN, int CIN, int CT>
__global__ void __launch_bounds__(kThreads)
smooth1_generic_kernel(SmArgs a, LossOut lo) {
  const int64_t hw = (int64_t)a.h * a.w;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  const int b = blockIdx.z;
  float loss = 0.f;
  // a block walks rows blockIdx.y, blockIdx.y + gridDim.y, ...: long-lived blocks keep the SM full (one-row
  // blocks spent their life being launched: 42 % achieved occupancy) and there is one loss hand-off per block
  if (x < a.w)
  for (int y = blockIdx.y; y < a.h; y += gridDim.y) {
    const int w = a.w, h = a.h;
    const float wx0 = w1_x<CIN, CT>(a, b, y, x), wy0 = w1_y<CIN, CT>(a, b, y, x);
    const bool need_g = a.grad != nullptr;
    const float wxm = (need_g && x > 0) ? w1_x<CIN, CT>(a, b, y, x - 1) : 0.f;
    const float wym = (need_g && y > 0) ? w1_y<CIN, CT>(a, b, y - 1, x) : 0.f;
    constexpr int NG = CIN ? CIN : 1;
    const int Cin = CIN ? CIN : a.Cin;
    float gsave[NG];
#pragma unroll
    for (int ch = 0; ch < Cin; ++ch) {
      const int64_t p = ((int64_t)b * Cin + ch) * hw + (int64_t)y * w + x;
      const float v = __ldg(a.in + p);
      const float gx = (x < w - 1) ? __ldg(a.in + p + 1) - v : 0.f;
      const float gy = (y < h - 1) ? __ldg(a.in + p + w) - v : 0.f;
      loss += pen_apply<PEN>(gx, a.eps2) * wx0 + pen_apply<PEN>(gy, a.eps2) * wy0;
      if (need_g) {
        // -Gx + shift(Gx) - Gy + shift(Gy)   (SmoothnessCriterion.lua:85-103)
        float g = -(pen_der<PEN>(gx, a.eps2) * wx0);
        if (x > 0) g += pen_der<PEN>(v - __ldg(a.in + p - 1), a.eps2) * wxm;
        g -= pen_der<PEN>(gy, a.eps2) * wy0;
        if (y > 0) g += pen_der<PEN>(v - __ldg(a.in + p - w), a.eps2) * wym;
        if (CIN) gsave[ch] = g * a.norm;
        else a.grad[p] = g * a.norm;
      }
    }
    if (CIN && need_g) {
#pragma unroll
      for (int ch = 0; ch < NG; ++ch) a.grad[((int64_t)b * Cin + ch) * hw + (int64_t)y * w + x] = gsave[ch];
    }
  }
  finish_loss(loss, lo);
}

// Model shapes (2-channel input, RGB target): a block owns `rows` CONSECUTIVE rows of a 128-pixel column strip.
// Every thread evaluates only its own two edge weights and G = p'(g) * w per row; the two shifted terms of the
// gradient, shiftR(Gx) and shiftD(Gy) (SmoothnessCriterion.lua:95-103), are the left neighbour's Gx (exchanged
// through shared memory; only thread 0 recomputes its outside neighbour) and the previous row's Gy (carried
// in registers; recomputed once at the top of the strip) -- the same operands and operations as evaluating them
// in place, so the result is bit-identical to the generic kernel, at half the weight / penalty evaluations and
// two instead of five input loads per channel.
template <int PEN>
__global__ void __launch_bounds__(kThreads)
smooth1_kernel(SmArgs a, LossOut lo, int rows) {
  constexpr int CIN = 2, CT = 3;
  __shared__ float s_gx[2][CIN][kThreads];
  const int64_t hw = (int64_t)a.h * a.w;
  const int w = a.w, h = a.h;
  const int tid = threadIdx.x;
  const int x = blockIdx.x * kThreads + tid;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * rows, y1 = min(h, y0 + rows);
  const bool act = x < w;
  const bool need_g = a.grad != nullptr;
  const float* in0 = a.in + ((int64_t)b * CIN) * hw + x;
  float loss = 0.f;
  float gy_up[CIN] = {0.f, 0.f}, v_cur[CIN] = {0.f, 0.f};
  if (act) {
#pragma unroll
    for (int ch = 0; ch < CIN; ++ch) v_cur[ch] = __ldg(in0 + ch * hw + (int64_t)y0 * w);
    if (need_g && y0 > 0) {
      const float wyu = w1_y<CIN, CT>(a, b, y0 - 1, x);
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch)
        gy_up[ch] = pen_der<PEN>(v_cur[ch] - __ldg(in0 + ch * hw + (int64_t)(y0 - 1) * w), a.eps2) * wyu;
    }
  }
  for (int y = y0; y < y1; ++y) {
    float Gx[CIN] = {0.f, 0.f}, Gy[CIN] = {0.f, 0.f}, vl[CIN] = {0.f, 0.f};
    float wxl = 0.f;
    if (act) {
      const float wx0 = w1_x<CIN, CT>(a, b, y, x), wy0 = w1_y<CIN, CT>(a, b, y, x);
      if (need_g && tid == 0 && x > 0) wxl = w1_x<CIN, CT>(a, b, y, x - 1);
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) {
        const float* p = in0 + ch * hw + (int64_t)y * w;
        const float v = v_cur[ch];
        const float vr = (x < w - 1) ? __ldg(p + 1) : 0.f;
        const float vd = (y < h - 1) ? __ldg(p + w) : 0.f;
        if (need_g && tid == 0 && x > 0) vl[ch] = __ldg(p - 1);
        const float gx = (x < w - 1) ? vr - v : 0.f;
        const float gy = (y < h - 1) ? vd - v : 0.f;
        loss += pen_apply<PEN>(gx, a.eps2) * wx0 + pen_apply<PEN>(gy, a.eps2) * wy0;
        Gx[ch] = pen_der<PEN>(gx, a.eps2) * wx0;
        Gy[ch] = pen_der<PEN>(gy, a.eps2) * wy0;
      }
    }
    if (need_g) {   // uniform for the block
      const int buf = y & 1;
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) s_gx[buf][ch][tid] = Gx[ch];
      __syncthreads();
      if (act) {
#pragma unroll
        for (int ch = 0; ch < CIN; ++ch) {
          // -Gx + shift(Gx) - Gy + shift(Gy), in the reference's order
          float g = -Gx[ch];
          if (x > 0) g += (tid > 0) ? s_gx[buf][ch][tid - 1] : pen_der<PEN>(v_cur[ch] - vl[ch], a.eps2) * wxl;
          g -= Gy[ch];
          if (y > 0) g += gy_up[ch];
          a.grad[((int64_t)b * CIN + ch) * hw + (int64_t)y * w + x] = g * a.norm;
        }
      }
    }
    if (act) {
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) {
        gy_up[ch] = Gy[ch];
        if (y < h - 1) v_cur[ch] = __ldg(in0 + ch * hw + (int64_t)(y + 1) * w);
      }
    }
  }
  finish_loss(loss, lo);
}

// Lean, software-pipelined form of smooth1_kernel for tensors whose flat indices fit 32 bits.  ncu put the kernel
// above at 0.46 IPC per scheduler with the memory system 7 % busy: ~600 issue slots per pixel (64-bit index
// arithmetic for every load, a lane-0-only path that the whole first warp steps through, per-thread recomputation
// of row-uniform quantities) and one dependent round trip per row.  Here
//   * thread 0 of a block is a halo column (x0 - 1): it only produces Gx for its right neighbour, so every thread
//     runs the same code (127 output columns per block);
//   * all offsets are 32-bit, row-uniform parts of the Q9 index arithmetic are hoisted out of the channel loops;
//   * the 16 loads of row y + 1 (input neighbours + the raw target pairs of both edge weights) are issued before
//     row y is evaluated, so two rows of loads are in flight per thread.
// Same operands and operation order per value as smooth1_kernel.
struct S1Raw {
  float vr[2], vd[2];           // input right / down neighbours
  float xa[3], xb[3];           // x-weight target pairs: |xb - xa| summed
  float ya[3], yb[3];           // y-weight target pairs
};

template <bool ALIAS>
__device__ __forceinline__ void s1_fetch(const SmArgs& a, const float* __restrict__ in0, unsigned b, int y, int x,
                                         bool inimg, S1Raw& r) {
  const unsigned w = a.w, h = a.h, hw = (unsigned)a.h * a.w;
#pragma unroll
  for (int i = 0; i < 3; ++i) r.xa[i] = r.xb[i] = r.ya[i] = r.yb[i] = 0.f;
  r.vr[0] = r.vr[1] = r.vd[0] = r.vd[1] = 0.f;
  if (!inimg) return;
  const unsigned off = (unsigned)y * w + x;
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    if ((unsigned)x < w - 1) r.vr[ch] = __ldg(in0 + ch * hw + off + 1);
    if ((unsigned)y < h - 1) r.vd[ch] = __ldg(in0 + ch * hw + off + w);
  }
  if (ALIAS) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const unsigned t = (b * 2 + j) * h + y;          // row counter of the (B,Cin,h,w) view
      const unsigned k = t * w + x;                      // flat index into the contiguous difference arrays
      if ((int64_t)k < a.n_dx) {
        const unsigned rowi = fast_div(k, w - 1, a.mul_w1, a.shr_w1);
        const float* p = a.tgt + rowi * w + (k - rowi * (w - 1));
        r.xa[j] = __ldg(p);
        r.xb[j] = __ldg(p + 1);
      }
      if ((int64_t)k < a.n_dy) {
        const unsigned t2 = fast_div(t, h - 1, a.mul_h1, a.shr_h1);
        const float* p = a.tgt + t2 * hw + (t - t2 * (h - 1)) * w + x;
        r.ya[j] = __ldg(p);
        r.yb[j] = __ldg(p + w);
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* p = a.tgt + (b * 3 + c) * hw + off;
      const float v = __ldg(p);
      r.xa[c] = r.ya[c] = v;
      r.xb[c] = ((unsigned)x < w - 1) ? __ldg(p + 1) : v;     // |v - v| = 0: the term is absent
      r.yb[c] = ((unsigned)y < h - 1) ? __ldg(p + w) : v;
    }
  }
}
template <bool ALIAS>
__device__ __forceinline__ float s1_weight(const float (&lo)[3], const float (&hi)[3], float cs) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) s += fabsf(hi[i] - lo[i]);     // the third alias term is |0 - 0|
  return __expf(-cs * (s / (ALIAS ? 2.f : 3.f)));
}

constexpr int kS1Out = kThreads - 1;
template <int PEN, bool ALIAS>
__global__ void __launch_bounds__(kThreads)
smooth1_lean_kernel(SmArgs a, LossOut lo, int rows) {
  constexpr int CIN = 2;
  __shared__ float s_gx[2][CIN][kThreads];
  const int w = a.w, h = a.h;
  const unsigned hw = (unsigned)h * w;
  const int tid = threadIdx.x;
  const int x = blockIdx.x * kS1Out + tid - 1;
  const unsigned b = blockIdx.z;
  const int y0 = blockIdx.y * rows, y1 = min(h, y0 + rows);
  const bool inimg = x >= 0 && x < w;
  const bool outp = inimg && tid >= 1;
  const bool need_g = a.grad != nullptr;
  const float* in0 = a.in + (size_t)b * CIN * hw;
  float loss = 0.f;
  float gy_up[CIN] = {0.f, 0.f}, v_cur[CIN] = {0.f, 0.f};
  S1Raw cur;
  s1_fetch<ALIAS>(a, in0, b, y0, x, inimg, cur);
  if (inimg) {
#pragma unroll
    for (int ch = 0; ch < CIN; ++ch) v_cur[ch] = __ldg(in0 + ch * hw + (unsigned)y0 * w + x);
    if (need_g && y0 > 0) {   // Gy of the row above the strip, once per strip
      S1Raw up;
      s1_fetch<ALIAS>(a, in0, b, y0 - 1, x, true, up);
      const float wyu = s1_weight<ALIAS>(up.ya, up.yb, a.cs);
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch)
        gy_up[ch] = pen_der<PEN>(v_cur[ch] - __ldg(in0 + ch * hw + (unsigned)(y0 - 1) * w + x), a.eps2) * wyu;
    }
  }
  for (int y = y0; y < y1; ++y) {
    S1Raw nxt;
    s1_fetch<ALIAS>(a, in0, b, y + 1, x, inimg && y + 1 < y1, nxt);   // in flight while row y is evaluated
    float Gx[CIN] = {0.f, 0.f}, Gy[CIN] = {0.f, 0.f};
    if (inimg) {
      const float wx0 = s1_weight<ALIAS>(cur.xa, cur.xb, a.cs), wy0 = s1_weight<ALIAS>(cur.ya, cur.yb, a.cs);
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) {
        const float v = v_cur[ch];
        const float gx = (x < w - 1) ? cur.vr[ch] - v : 0.f;
        const float gy = (y < h - 1) ? cur.vd[ch] - v : 0.f;
        if (outp) loss += pen_apply<PEN>(gx, a.eps2) * wx0 + pen_apply<PEN>(gy, a.eps2) * wy0;
        Gx[ch] = pen_der<PEN>(gx, a.eps2) * wx0;
        Gy[ch] = pen_der<PEN>(gy, a.eps2) * wy0;
      }
    }
    if (need_g) {   // uniform for the block
      const int buf = y & 1;
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) s_gx[buf][ch][tid] = Gx[ch];
      __syncthreads();
      if (outp) {
#pragma unroll
        for (int ch = 0; ch < CIN; ++ch) {
          // -Gx + shift(Gx) - Gy + shift(Gy), in the reference's order (SmoothnessCriterion.lua:95-103)
          float g = -Gx[ch];
          if (x > 0) g += s_gx[buf][ch][tid - 1];
          g -= Gy[ch];
          if (y > 0) g += gy_up[ch];
          a.grad[(size_t)b * CIN * hw + ch * hw + (unsigned)y * w + x] = g * a.norm;
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < CIN; ++ch) {
      gy_up[ch] = Gy[ch];
      v_cur[ch] = cur.vd[ch];   // the down neighbour is the next row's centre value
    }
    cur = nxt;
  }
  finish_loss(loss, lo);
}

// order-2 weights (SecondOrderSmoothnessCriterion.lua:49-61) ------------------------------
template <int CIN, int CT>
__device__ __forceinline__ float w2_y(const SmArgs& a, int b, int y, int x) {
  const int64_t hw = (int64_t)a.h * a.w;
  const int Cin = CIN ? CIN : a.Cin, Ct = CT ? CT : a.Ct;
  (void)Cin; (void)Ct;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < Ct; ++c) {
    const float* p = a.tgt + ((int64_t)b * Ct + c) * hw + (int64_t)y * a.w + x;
    const float t = __ldg(p);
    if (y >= 1) s1 += fabsf(t - __ldg(p - a.w));
    if (y >= 1 && y <= a.h - 2) s2 += fabsf(t - __ldg(p + a.w));
  }
  return __expf(-a.cs * (s1 / (float)Ct + s2 / (float)Ct));
}
template <int CIN, int CT>
__device__ __forceinline__ float w2_x(const SmArgs& a, int b, int y, int x) {
  const int64_t hw = (int64_t)a.h * a.w;
  const int Cin = CIN ? CIN : a.Cin, Ct = CT ? CT : a.Ct;
  (void)Cin; (void)Ct;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < Ct; ++c) {
    const float* p = a.tgt + ((int64_t)b * Ct + c) * hw + (int64_t)y * a.w + x;
    const float t = __ldg(p);
    if (x >= 1) s1 += fabsf(t - __ldg(p - 1));
    if (x >= 1 && x <= a.w - 2) s2 += fabsf(t - __ldg(p + 1));
  }
  return __expf(-a.cs * (s1 / (float)Ct + s2 / (float)Ct));
}

template <int PEN, int CIN, int CT>
__global__ void __launch_bounds__(kThreads)
smooth2_generic_kernel(SmArgs a, LossOut lo) {
  const int64_t hw = (int64_t)a.h * a.w;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  const int b = blockIdx.z;
  float loss = 0.f;
  // a block walks rows blockIdx.y, blockIdx.y + gridDim.y, ...: long-lived blocks keep the SM full (one-row
  // blocks spent their life being launched: 42 % achieved occupancy) and there is one loss hand-off per block
  if (x < a.w)
  for (int y = blockIdx.y; y < a.h; y += gridDim.y) {
    const int w = a.w, h = a.h;
    const bool need_g = a.grad != nullptr;
    // weights at the three positions each direction needs
    float wy[3], wx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int yy = y + k - 1, xx = x + k - 1;
      wy[k] = (yy >= 0 && yy < h && (k == 1 || need_g)) ? w2_y<CIN, CT>(a, b, yy, x) : 0.f;
      wx[k] = (xx >= 0 && xx < w && (k == 1 || need_g)) ? w2_x<CIN, CT>(a, b, y, xx) : 0.f;
    }
    constexpr int NG = CIN ? CIN : 1;
    const int Cin = CIN ? CIN : a.Cin;
    float gsave[NG];
#pragma unroll
    for (int ch = 0; ch < Cin; ++ch) {
      const float* I = a.in + ((int64_t)b * Cin + ch) * hw;
      const int64_t o = (int64_t)y * w + x;
      // second differences at y-1, y, y+1 (zero outside the interior 1..h-2)
      float gy[3], gx[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int yy = y + k - 1, xx = x + k - 1;
        gy[k] = 0.f;
        gx[k] = 0.f;
        if ((k == 1 || need_g) && yy >= 1 && yy <= h - 2) {
          const int64_t q = (int64_t)yy * w + x;
          gy[k] = (2.f * __ldg(I + q) - __ldg(I + q - w)) - __ldg(I + q + w);
        }
        if ((k == 1 || need_g) && xx >= 1 && xx <= w - 2) {
          const int64_t q = (int64_t)y * w + xx;
          gx[k] = (2.f * __ldg(I + q) - __ldg(I + q - 1)) - __ldg(I + q + 1);
        }
      }
      loss += pen_apply<PEN>(gx[1], a.eps2) * wx[1] + pen_apply<PEN>(gy[1], a.eps2) * wy[1];
      if (need_g) {
        // adjoint of the second difference (SecondOrderSmoothnessCriterion.lua:90-97)
        float g = 0.f;
        if (y >= 1 && y <= h - 2) g += 2.f * (pen_der<PEN>(gy[1], a.eps2) * wy[1]);
        if (x >= 1 && x <= w - 2) g += 2.f * (pen_der<PEN>(gx[1], a.eps2) * wx[1]);
        if (y + 1 >= 1 && y + 1 <= h - 2) g -= pen_der<PEN>(gy[2], a.eps2) * wy[2];
        if (x + 1 >= 1 && x + 1 <= w - 2) g -= pen_der<PEN>(gx[2], a.eps2) * wx[2];
        if (y - 1 >= 1 && y - 1 <= h - 2) g -= pen_der<PEN>(gy[0], a.eps2) * wy[0];
        if (x - 1 >= 1 && x - 1 <= w - 2) g -= pen_der<PEN>(gx[0], a.eps2) * wx[0];
        if (CIN) gsave[ch] = g * a.norm;
        else a.grad[((int64_t)b * Cin + ch) * hw + o] = g * a.norm;
      }
    }
    if (CIN && need_g) {
#pragma unroll
      for (int ch = 0; ch < NG; ++ch) a.grad[((int64_t)b * Cin + ch) * hw + (int64_t)y * w + x] = gsave[ch];
    }
  }
  finish_loss(loss, lo);
}

// Model shapes (2-channel input, RGB target).  grad = 2 G[i] - G[i-1] - G[i+1] per direction with G = p'(g) * w
// (SecondOrderSmoothnessCriterion.lua:87-97): the generic kernel evaluates G at all six neighbours of a pixel.
// Here a block owns `rows` consecutive rows of a 126-pixel column strip (threads 0 and 127 are halo pixels):
// every thread evaluates G once per row, Gx of the neighbours comes through shared memory, Gy is computed one
// row ahead and kept in a three-row register window, and the input rows slide through registers (one new load
// per channel and row).  Same operands; the six terms are summed in a different order (fp32 rounding only).
constexpr int kS2Out = kThreads - 2;
template <int PEN>
__global__ void __launch_bounds__(kThreads)
smooth2_kernel(SmArgs a, LossOut lo, int rows) {
  constexpr int CIN = 2, CT = 3;
  __shared__ float s_gx[2][CIN][kThreads];
  const int64_t hw = (int64_t)a.h * a.w;
  const int w = a.w, h = a.h;
  const int tid = threadIdx.x;
  const int x = blockIdx.x * kS2Out + tid - 1;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * rows, y1 = min(h, y0 + rows);
  const bool inimg = x >= 0 && x < w;
  const bool outp = inimg && tid >= 1 && tid <= kS2Out;
  const bool need_g = a.grad != nullptr;
  const float* in0 = a.in + ((int64_t)b * CIN) * hw + (inimg ? x : 0);
  float loss = 0.f;
  // register window over the input rows r-1, r, r+1 and over Gy(r-2), Gy(r-1)
  float im1[CIN], i0[CIN], ip1[CIN], gy2[CIN], gy1[CIN], xterm[CIN], gyc[CIN];
#pragma unroll
  for (int ch = 0; ch < CIN; ++ch) {
    im1[ch] = i0[ch] = ip1[ch] = gy2[ch] = gy1[ch] = xterm[ch] = gyc[ch] = 0.f;
    if (inimg) {
      const int r = y0 - 1;   // first row whose Gy is needed
      if (r - 1 >= 0) im1[ch] = __ldg(in0 + ch * hw + (int64_t)(r - 1) * w);
      if (r >= 0) i0[ch] = __ldg(in0 + ch * hw + (int64_t)r * w);
      if (r + 1 < h) ip1[ch] = __ldg(in0 + ch * hw + (int64_t)(r + 1) * w);
    }
  }
  const int rbeg = need_g ? y0 - 1 : y0, rend = need_g ? y1 : y1 - 1;
  if (!need_g && inimg) {   // the window starts one row later
#pragma unroll
    for (int ch = 0; ch < CIN; ++ch) {
      im1[ch] = i0[ch];
      i0[ch] = ip1[ch];
      ip1[ch] = (y0 + 1 < h) ? __ldg(in0 + ch * hw + (int64_t)(y0 + 1) * w) : 0.f;
    }
  }
  for (int r = rbeg; r <= rend; ++r) {
    const bool own = r >= y0 && r < y1;   // a row of this block (loss, x terms); else only Gy is needed
    float Gx[CIN] = {0.f, 0.f};
    if (inimg && r >= 0 && r < h) {
      const float wy = w2_y<CIN, CT>(a, b, r, x);
      const float wx = own ? w2_x<CIN, CT>(a, b, r, x) : 0.f;
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) {
        const float gy = (r >= 1 && r <= h - 2) ? (2.f * i0[ch] - im1[ch]) - ip1[ch] : 0.f;
        gyc[ch] = (r >= 1 && r <= h - 2) ? pen_der<PEN>(gy, a.eps2) * wy : 0.f;
        if (own) {
          float gx = 0.f;
          if (x >= 1 && x <= w - 2) {
            const float* q = in0 + ch * hw + (int64_t)r * w;
            gx = (2.f * i0[ch] - __ldg(q - 1)) - __ldg(q + 1);
            Gx[ch] = pen_der<PEN>(gx, a.eps2) * wx;
          }
          if (outp) loss += pen_apply<PEN>(gx, a.eps2) * wx + pen_apply<PEN>(gy, a.eps2) * wy;
        }
      }
    } else {
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) gyc[ch] = 0.f;
    }
    if (need_g) {   // uniform for the block
      const int buf = r & 1;
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) s_gx[buf][ch][tid] = Gx[ch];
      __syncthreads();
      if (outp) {
#pragma unroll
        for (int ch = 0; ch < CIN; ++ch) {
          // row r-1 is complete: 2 Gy(r-1) - Gy(r) - Gy(r-2) + its x terms
          if (r - 1 >= y0) {
            const float g = ((2.f * gy1[ch] - gyc[ch]) - gy2[ch]) + xterm[ch];
            a.grad[((int64_t)b * CIN + ch) * hw + (int64_t)(r - 1) * w + x] = g * a.norm;
          }
          if (own) xterm[ch] = (2.f * Gx[ch] - s_gx[buf][ch][tid + 1]) - s_gx[buf][ch][tid - 1];
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < CIN; ++ch) {
      gy2[ch] = gy1[ch];
      gy1[ch] = gyc[ch];
      im1[ch] = i0[ch];
      i0[ch] = ip1[ch];
      ip1[ch] = (inimg && r + 2 < h && r + 2 >= 0) ? __ldg(in0 + ch * hw + (int64_t)(r + 2) * w) : 0.f;
    }
  }
  finish_loss(loss, lo);
}

// Lean, software-pipelined form of smooth2_kernel (32-bit offsets; see smooth1_lean_kernel).  Row r needs the target
// at rows r-1, r, r+1 and at x-1, x+1, and the input at x-1, x+1 and rows r-1, r, r+1.  Centre columns slide
// through three-row register windows (one new target row and one new input row per iteration); everything
// iteration r + 1 needs from memory is requested before iteration r is evaluated.
struct S2Raw {
  float tc2[3];          // target, centre column, row r + 1
  float tl[3], tr[3];    // target, row r, x - 1 / x + 1
  float xl[2], xr[2];    // input, row r, x - 1 / x + 1
  float in2[2];          // input, centre column, row r + 1
};
__device__ __forceinline__ void s2_fetch(const SmArgs& a, const float* __restrict__ in0, const float* __restrict__ t0,
                                         int r, int x, bool on, bool own, S2Raw& q) {
  const int w = a.w, h = a.h;
  const unsigned hw = (unsigned)h * w;
#pragma unroll
  for (int c = 0; c < 3; ++c) q.tc2[c] = q.tl[c] = q.tr[c] = 0.f;
  q.xl[0] = q.xl[1] = q.xr[0] = q.xr[1] = q.in2[0] = q.in2[1] = 0.f;
  if (!on) return;
  const bool row_ok = r >= 0 && r < h;
  const unsigned off = (unsigned)(row_ok ? r : 0) * w + x;
  if (r + 1 >= 0 && r + 1 < h) {
#pragma unroll
    for (int c = 0; c < 3; ++c) q.tc2[c] = __ldg(t0 + c * hw + (unsigned)(r + 1) * w + x);
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) q.in2[ch] = __ldg(in0 + ch * hw + (unsigned)(r + 1) * w + x);
  }
  if (row_ok && own) {
    if (x >= 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) q.tl[c] = __ldg(t0 + c * hw + off - 1);
    }
    if (x >= 1 && x <= w - 2) {
#pragma unroll
      for (int c = 0; c < 3; ++c) q.tr[c] = __ldg(t0 + c * hw + off + 1);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        q.xl[ch] = __ldg(in0 + ch * hw + off - 1);
        q.xr[ch] = __ldg(in0 + ch * hw + off + 1);
      }
    }
  }
}

template <int PEN>
__global__ void __launch_bounds__(kThreads)
smooth2_lean_kernel(SmArgs a, LossOut lo, int rows) {
  constexpr int CIN = 2, CT = 3;
  __shared__ float s_gx[2][CIN][kThreads];
  const int w = a.w, h = a.h;
  const unsigned hw = (unsigned)h * w;
  const int tid = threadIdx.x;
  const int x = blockIdx.x * kS2Out + tid - 1;
  const unsigned b = blockIdx.z;
  const int y0 = blockIdx.y * rows, y1 = min(h, y0 + rows);
  const bool inimg = x >= 0 && x < w;
  const bool outp = inimg && tid >= 1 && tid <= kS2Out;
  const bool need_g = a.grad != nullptr;
  const float* in0 = a.in + (size_t)b * CIN * hw;
  const float* t0 = a.tgt + (size_t)b * CT * hw;
  float loss = 0.f;
  const int rbeg = need_g ? y0 - 1 : y0, rend = need_g ? y1 : y1 - 1;
  // windows over rows r-1, r (r+1 arrives with the prefetched data) and over Gy(r-2), Gy(r-1)
  float im1[CIN], i0[CIN], tm1[CT], tc[CT], gy2[CIN], gy1[CIN], xterm[CIN], gyc[CIN];
#pragma unroll
  for (int ch = 0; ch < CIN; ++ch) {
    im1[ch] = i0[ch] = gy2[ch] = gy1[ch] = xterm[ch] = gyc[ch] = 0.f;
    if (inimg) {
      if (rbeg - 1 >= 0) im1[ch] = __ldg(in0 + ch * hw + (unsigned)(rbeg - 1) * w + x);
      if (rbeg >= 0) i0[ch] = __ldg(in0 + ch * hw + (unsigned)rbeg * w + x);
    }
  }
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    tm1[c] = tc[c] = 0.f;
    if (inimg) {
      if (rbeg - 1 >= 0) tm1[c] = __ldg(t0 + c * hw + (unsigned)(rbeg - 1) * w + x);
      if (rbeg >= 0) tc[c] = __ldg(t0 + c * hw + (unsigned)rbeg * w + x);
    }
  }
  S2Raw cur;
  s2_fetch(a, in0, t0, rbeg, x, inimg, rbeg >= y0 && rbeg < y1, cur);
  for (int r = rbeg; r <= rend; ++r) {
    const bool own = r >= y0 && r < y1;   // a row of this block (loss, x terms); else only Gy is needed
    S2Raw nxt;
    s2_fetch(a, in0, t0, r + 1, x, inimg && r + 1 <= rend, r + 1 >= y0 && r + 1 < y1, nxt);
    float Gx[CIN] = {0.f, 0.f};
    if (inimg && r >= 0 && r < h) {
      // edge weights (SecondOrderSmoothnessCriterion.lua:49-61), same accumulation order as w2_y / w2_x
      float sy1 = 0.f, sy2 = 0.f, sx1 = 0.f, sx2 = 0.f;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        if (r >= 1) sy1 += fabsf(tc[c] - tm1[c]);
        if (r >= 1 && r <= h - 2) sy2 += fabsf(tc[c] - cur.tc2[c]);
        if (x >= 1) sx1 += fabsf(tc[c] - cur.tl[c]);
        if (x >= 1 && x <= w - 2) sx2 += fabsf(tc[c] - cur.tr[c]);
      }
      const float wy = __expf(-a.cs * (sy1 / (float)CT + sy2 / (float)CT));
      const float wx = own ? __expf(-a.cs * (sx1 / (float)CT + sx2 / (float)CT)) : 0.f;
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) {
        const float gy = (r >= 1 && r <= h - 2) ? (2.f * i0[ch] - im1[ch]) - cur.in2[ch] : 0.f;
        gyc[ch] = (r >= 1 && r <= h - 2) ? pen_der<PEN>(gy, a.eps2) * wy : 0.f;
        if (own) {
          float gx = 0.f;
          if (x >= 1 && x <= w - 2) {
            gx = (2.f * i0[ch] - cur.xl[ch]) - cur.xr[ch];
            Gx[ch] = pen_der<PEN>(gx, a.eps2) * wx;
          }
          if (outp) loss += pen_apply<PEN>(gx, a.eps2) * wx + pen_apply<PEN>(gy, a.eps2) * wy;
        }
      }
    } else {
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) gyc[ch] = 0.f;
    }
    if (need_g) {   // uniform for the block
      const int buf = r & 1;
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) s_gx[buf][ch][tid] = Gx[ch];
      __syncthreads();
      if (outp) {
#pragma unroll
        for (int ch = 0; ch < CIN; ++ch) {
          // row r-1 is complete: 2 Gy(r-1) - Gy(r) - Gy(r-2) + its x terms
          if (r - 1 >= y0) {
            const float g = ((2.f * gy1[ch] - gyc[ch]) - gy2[ch]) + xterm[ch];
            a.grad[(size_t)b * CIN * hw + ch * hw + (unsigned)(r - 1) * w + x] = g * a.norm;
          }
          if (own) xterm[ch] = (2.f * Gx[ch] - s_gx[buf][ch][tid + 1]) - s_gx[buf][ch][tid - 1];
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < CIN; ++ch) {
      gy2[ch] = gy1[ch];
      gy1[ch] = gyc[ch];
      im1[ch] = i0[ch];
      i0[ch] = cur.in2[ch];
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      tm1[c] = tc[c];
      tc[c] = cur.tc2[c];
    }
    cur = nxt;
  }
  finish_loss(loss, lo);
}

// =======================================================================================
// ConstVelCriterion / OcclusionPriorCriterion
// =======================================================================================
__global__ void __launch_bounds__(kThreads)
constvel_kernel(const float* __restrict__ f, const float* __restrict__ bb, float* __restrict__ gf,
                float* __restrict__ gb, int B, int C, int64_t hw, float gnorm, LossOut lo) {
  const int b = blockIdx.y;
  float loss = 0.f;
  // pixel inside the image; a block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
  for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < hw; o += (int64_t)gridDim.x * kThreads) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const int64_t p = ((int64_t)b * C + c) * hw + o;
      const float d = __ldg(f + p) - __ldg(bb + p);
      s += d * d;
    }
    const float nrm = sqrtf(s);
    loss += nrm;
    if (gf || gb) {
      const float den = nrm + 1e-12f;
      for (int c = 0; c < C; ++c) {
        const int64_t p = ((int64_t)b * C + c) * hw + o;
        const float fv = __ldg(f + p), bv = __ldg(bb + p);
        if (gf) gf[p] = ((fv - bv) / den) * gnorm;
        if (gb) gb[p] = ((bv - fv) / den) * gnorm;
      }
    }
  }
  finish_loss(loss, lo);
}

__global__ void __launch_bounds__(kThreads)
occprior_kernel(const float* __restrict__ occ, float* __restrict__ grad, int B, int C, int64_t hw,
                float penalty, float norm, LossOut lo) {
  const int b = blockIdx.y;
  float loss = 0.f;
  for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < hw; o += (int64_t)gridDim.x * kThreads) {
    const int64_t p0 = ((int64_t)b * C) * hw + o;
    const float o1 = __ldg(occ + p0), o2 = __ldg(occ + p0 + hw);
    if (C == 3) {
      const float o3 = __ldg(occ + p0 + 2 * hw);
      loss += (1.f - o2) * (o1 + o3) * penalty * 0.05f;
      if (grad) {
        grad[p0] = (1.f - o2) * penalty * 0.05f * norm;
        grad[p0 + hw] = -(o1 + o3) * penalty * 0.05f * norm;
        grad[p0 + 2 * hw] = (1.f - o2) * penalty * 0.05f * norm;
      }
    } else {
      loss += (1.f - o1 * o2) * penalty;
      if (grad) {
        grad[p0] = (1.f - o2) * penalty * norm;
        grad[p0 + hw] = (1.f - o1) * penalty * norm;
      }
    }
  }
  finish_loss(loss, lo);
}

// Model shapes (two flow / occlusion channels, hw % 4 == 0, 16-byte aligned planes): four pixels per thread,
// every plane read and written with 128-bit accesses and all loads of the four pixels issued before the first
// use.  The scalar kernels above keep 4 bytes per load in flight per thread, far too little for the ~6 MB the
// HBM system needs in flight (1.7 TB/s measured); this form is what a streaming kernel has to look like here.
__device__ __forceinline__ float4 ldg_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__global__ void __launch_bounds__(kThreads)
constvel_c2_kernel(const float* __restrict__ f, const float* __restrict__ bb, float* __restrict__ gf,
                   float* __restrict__ gb, int64_t hw, float gnorm, LossOut lo) {
  const int b = blockIdx.y;
  const float* f0 = f + (int64_t)b * 2 * hw;
  const float* b0 = bb + (int64_t)b * 2 * hw;
  float loss = 0.f;
  const int64_t nq = hw >> 2;
  for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < nq; q += (int64_t)gridDim.x * kThreads) {
    const int64_t o = q << 2;
    const float4 fu = ldg_stream4(f0 + o), fv = ldg_stream4(f0 + hw + o);
    const float4 bu = ldg_stream4(b0 + o), bv = ldg_stream4(b0 + hw + o);
    const float du[4] = {fu.x - bu.x, fu.y - bu.y, fu.z - bu.z, fu.w - bu.w};
    const float dv[4] = {fv.x - bv.x, fv.y - bv.y, fv.z - bv.z, fv.w - bv.w};
    float gu[4], gv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s = 0.f;            // same accumulation order as the channel loop of the generic kernel
      s += du[i] * du[i];
      s += dv[i] * dv[i];
      const float nrm = sqrtf(s);
      loss += nrm;
      const float den = nrm + 1e-12f;
      gu[i] = (du[i] / den) * gnorm;
      gv[i] = (dv[i] / den) * gnorm;
    }
    if (gf) {
      stg_stream4(gf + (int64_t)b * 2 * hw + o, make_float4(gu[0], gu[1], gu[2], gu[3]));
      stg_stream4(gf + (int64_t)b * 2 * hw + hw + o, make_float4(gv[0], gv[1], gv[2], gv[3]));
    }
    if (gb) {   // (b - f) / den = -((f - b) / den) exactly
      stg_stream4(gb + (int64_t)b * 2 * hw + o, make_float4(-gu[0], -gu[1], -gu[2], -gu[3]));
      stg_stream4(gb + (int64_t)b * 2 * hw + hw + o, make_float4(-gv[0], -gv[1], -gv[2], -gv[3]));
    }
  }
  finish_loss(loss, lo);
}

__global__ void __launch_bounds__(kThreads)
occprior_c2_kernel(const float* __restrict__ occ, float* __restrict__ grad, int64_t hw, float penalty, float norm,
                   LossOut lo) {
  const int b = blockIdx.y;
  const float* o0 = occ + (int64_t)b * 2 * hw;
  float loss = 0.f;
  const int64_t nq = hw >> 2;
  for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < nq; q += (int64_t)gridDim.x * kThreads) {
    const int64_t o = q << 2;
    const float4 a = ldg_stream4(o0 + o), c = ldg_stream4(o0 + hw + o);
    loss += (1.f - a.x * c.x) * penalty;
    loss += (1.f - a.y * c.y) * penalty;
    loss += (1.f - a.z * c.z) * penalty;
    loss += (1.f - a.w * c.w) * penalty;
    if (grad) {
      float* g0 = grad + (int64_t)b * 2 * hw + o;
      stg_stream4(g0, make_float4((1.f - c.x) * penalty * norm, (1.f - c.y) * penalty * norm,
                                  (1.f - c.z) * penalty * norm, (1.f - c.w) * penalty * norm));
      stg_stream4(g0 + hw, make_float4((1.f - a.x) * penalty * norm, (1.f - a.y) * penalty * norm,
                                       (1.f - a.z) * penalty * norm, (1.f - a.w) * penalty * norm));
    }
  }
  finish_loss(loss, lo);
}

bool planes16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Number of row (or tile) slots in the grid when `cols` blocks exist per slot and there are `n` rows: about one
// wave of resident blocks (16 x 128 threads per SM), at most 16 rows per block.
int rows_per_grid(int64_t cols, int n) {
  const int64_t want = (int64_t)num_sms() * 16;
  int64_t gy = (want + cols - 1) / cols;
  const int64_t gmin = (n + 15) / 16;
  if (gy < gmin) gy = gmin;
  if (gy > n) gy = n;
  return (int)(gy < 1 ? 1 : gy);
}

// grid = (x tiles, rows, batch) for the stencil kernels
int grid_rows(int B, int h, int w, dim3* g, int* blocks) {
  if (B > 65535 || h > 65535) return fail(B2F_EINVAL, "criterion: B and h must be <= 65535");
  const int gx = (w + kThreads - 1) / kThreads;
  *g = dim3(gx, rows_per_grid((int64_t)gx * B, h), B);
  const int64_t nb = (int64_t)g->x * g->y * g->z;
  if (nb > 0x3fffffff) return fail(B2F_EINVAL, "criterion: too many pixels");
  *blocks = (int)nb;
  return B2F_OK;
}
// grid = (pixel tiles, batch) for the pointwise kernels
int grid_flat(int B, int64_t hw, dim3* g, int* blocks) {
  if (B > 65535) return fail(B2F_EINVAL, "criterion: B must be <= 65535");
  int64_t gx = (hw + kThreads - 1) / kThreads;
  if (gx > 0x3fffffff) return fail(B2F_EINVAL, "criterion: too many pixels");
  gx = rows_per_grid(B, (int)gx);
  *g = dim3((unsigned)gx, B, 1);
  *blocks = (int)(gx * B);
  return B2F_OK;
}

// B2F_SMOOTH_LEGACY=1 selects the older fast kernels (experiments / A-B timing)
bool smooth_legacy() {
  static const bool v = [] { const char* e = getenv("B2F_SMOOTH_LEGACY"); return e && e[0] == '1'; }();
  return v;
}

bool valid_penalty(int p) { return p == B2F_PENALTY_QUADRATIC || p == B2F_PENALTY_L1 || p == B2F_PENALTY_LORENTZIAN; }

template <bool GT>
void launch_ob(int pen, dim3 blocks, cudaStream_t st, const ObArgs& a, const LossOut& lo) {
  if (pen == B2F_PENALTY_QUADRATIC) ob_kernel<B2F_PENALTY_QUADRATIC, GT><<<blocks, kThreads, 0, st>>>(a, lo);
  else if (pen == B2F_PENALTY_L1) ob_kernel<B2F_PENALTY_L1, GT><<<blocks, kThreads, 0, st>>>(a, lo);
  else ob_kernel<B2F_PENALTY_LORENTZIAN, GT><<<blocks, kThreads, 0, st>>>(a, lo);
}

}  // namespace
}  // namespace b2f

using namespace b2f;

extern "C" int b2f_ob_criterion(const b2f_ob_params* prm, const float* flow, const float* bflow,
                                const float* occ, const float* warp_past, const float* warp_future,
                                const float* target, int B, int C, int h, int w, float* grad_occ,
                                float* grad_warp_past, float* grad_warp_future, double* loss_dev,
                                double* loss_host, b2f_stream_t stream) {
  if (!prm) return fail(B2F_EINVAL, "ob_criterion: params is NULL");
  if (!flow || !occ || !warp_past || !warp_future || !target) return fail(B2F_EINVAL, "ob_criterion: NULL input");
  if (prm->past_flow && !bflow) return fail(B2F_EINVAL, "ob_criterion: past_flow set but bflow is NULL");
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return fail(B2F_EINVAL, "ob_criterion: bad size B=%d C=%d h=%d w=%d", B, C, h, w);
  if (!valid_penalty(prm->penalty)) return fail(B2F_EINVAL, "ob_criterion: unknown penalty %d", prm->penalty);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t npix = (int64_t)B * h * w;
  int blocks;
  dim3 grid;
  int rc = grid_rows(B, h, w, &grid, &blocks);
  if (rc) return rc;
  const int F = 3;
  double scale = 1.0 / ((double)C * (F - 1));
  if (prm->size_average) scale *= 1.0 / (double)npix;
  ObArgs a;
  a.flow = flow;
  a.bflow = prm->past_flow ? bflow : flow;
  a.occ = occ;
  a.warp[0] = warp_past;
  a.warp[1] = warp_future;
  a.target = target;
  a.g_occ = grad_occ;
  a.g_warp[0] = grad_warp_past;
  a.g_warp[1] = grad_warp_future;
  a.B = B; a.C = C; a.h = h; a.w = w;
  a.eps2 = prm->penalty_eps > 0.f ? prm->penalty_eps * prm->penalty_eps : 0.05f * 0.05f;
  a.penalty_out = prm->penalty_out;
  a.alpha = prm->alpha; a.beta = prm->beta; a.gamma = prm->gamma;
  a.scale = prm->pwc_flow_scaling;
  a.norm = (float)scale;
  a.grad_check = prm->grad_check;
  LossScratch ls;
  if ((rc = ls.begin(blocks, scale, loss_dev, st))) return rc;
  if (prm->gradient_terms) launch_ob<true>(prm->penalty, grid, st, a, ls.lo);
  else launch_ob<false>(prm->penalty, grid, st, a, ls.lo);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) count_launch();
  rc = ls.end(loss_host);
  if (e != cudaSuccess) return cuda_fail(e, "ob_kernel");
  return rc;
}

extern "C" int b2f_smoothness_criterion(const b2f_smooth_params* prm, const float* input, const float* target,
                                        int B, int Cin, int Ct, int h, int w, float* grad, double* loss_dev,
                                        double* loss_host, b2f_stream_t stream) {
  if (!prm) return fail(B2F_EINVAL, "smoothness: params is NULL");
  if (!input || !target) return fail(B2F_EINVAL, "smoothness: NULL input/target");
  if (B <= 0 || Cin <= 0 || Ct <= 0 || h <= 0 || w <= 0) return fail(B2F_EINVAL, "smoothness: bad size");
  if (prm->order != 1 && prm->order != 2) return fail(B2F_EINVAL, "smoothness: order %d", prm->order);
  if (!valid_penalty(prm->penalty)) return fail(B2F_EINVAL, "smoothness: unknown penalty %d", prm->penalty);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t npix = (int64_t)B * h * w;
  int blocks;
  dim3 grid;
  int rc = grid_rows(B, h, w, &grid, &blocks);
  if (rc) return rc;
  const double scale = prm->size_average ? 1.0 / ((double)npix * Cin) : 1.0;
  SmArgs a;
  a.in = input; a.tgt = target; a.grad = grad;
  a.B = B; a.Cin = Cin; a.Ct = Ct; a.h = h; a.w = w;
  a.eps2 = prm->penalty_eps > 0.f ? prm->penalty_eps * prm->penalty_eps : 0.05f * 0.05f;
  a.cs = prm->cs;
  a.norm = (float)scale;
  a.alias = (prm->alias_weights && Cin != Ct) ? 1 : 0;
  a.n_dy = (int64_t)B * Ct * (h - 1) * w;
  a.n_dx = (int64_t)B * Ct * h * (w - 1);
  a.small = ((int64_t)B * (Cin > Ct ? Cin : Ct) * h * w) < ((int64_t)1 << 31) ? 1 : 0;
  auto find_divisor = [](uint32_t d, uint32_t* mul, uint32_t* shr) {
    *mul = 0;
    *shr = 0;
    if (d > 1) {
      uint32_t lg = 0;
      while ((1ull << lg) < d) ++lg;
      const uint32_t pw = 31 + lg;
      *mul = (uint32_t)((((uint64_t)1 << pw) + d - 1) / d);
      *shr = pw - 32;
    }
  };
  find_divisor(h > 1 ? (uint32_t)(h - 1) : 1u, &a.mul_h1, &a.shr_h1);
  find_divisor(w > 1 ? (uint32_t)(w - 1) : 1u, &a.mul_w1, &a.shr_w1);
  LossScratch ls;
  // the second-order fast kernel uses 126-pixel x tiles: size the partials for the larger grid
  blocks = (int)((int64_t)((w + kS2Out - 1) / kS2Out) * grid.y * grid.z);
  if ((rc = ls.begin(blocks, scale, loss_dev, st))) return rc;
  const int pen = prm->penalty;
  const bool fixed = Cin == 2 && Ct == 3;   // the model's shapes: flow / occlusion map vs RGB target
  if (prm->order == 1 && fixed && a.small && !smooth_legacy()) {
    const int rows = (h + (int)grid.y - 1) / (int)grid.y;   // consecutive rows per block (measured flat from 4 to 14)
    const dim3 g1((w + kS1Out - 1) / kS1Out, (h + rows - 1) / rows, grid.z);
#define B2F_S1(P) do { if (a.alias) smooth1_lean_kernel<P, true><<<g1, kThreads, 0, st>>>(a, ls.lo, rows); \
                       else smooth1_lean_kernel<P, false><<<g1, kThreads, 0, st>>>(a, ls.lo, rows); } while (0)
    if (pen == B2F_PENALTY_QUADRATIC) B2F_S1(B2F_PENALTY_QUADRATIC);
    else if (pen == B2F_PENALTY_L1) B2F_S1(B2F_PENALTY_L1);
    else B2F_S1(B2F_PENALTY_LORENTZIAN);
#undef B2F_S1
  } else if (prm->order == 1 && fixed) {
    const int rows = (h + (int)grid.y - 1) / (int)grid.y;   // consecutive rows per block
    const dim3 g1(grid.x, (h + rows - 1) / rows, grid.z);
    if (pen == B2F_PENALTY_QUADRATIC) smooth1_kernel<B2F_PENALTY_QUADRATIC><<<g1, kThreads, 0, st>>>(a, ls.lo, rows);
    else if (pen == B2F_PENALTY_L1) smooth1_kernel<B2F_PENALTY_L1><<<g1, kThreads, 0, st>>>(a, ls.lo, rows);
    else smooth1_kernel<B2F_PENALTY_LORENTZIAN><<<g1, kThreads, 0, st>>>(a, ls.lo, rows);
  } else if (prm->order == 1) {
    if (pen == B2F_PENALTY_QUADRATIC) smooth1_generic_kernel<B2F_PENALTY_QUADRATIC, 0, 0><<<grid, kThreads, 0, st>>>(a, ls.lo);
    else if (pen == B2F_PENALTY_L1) smooth1_generic_kernel<B2F_PENALTY_L1, 0, 0><<<grid, kThreads, 0, st>>>(a, ls.lo);
    else smooth1_generic_kernel<B2F_PENALTY_LORENTZIAN, 0, 0><<<grid, kThreads, 0, st>>>(a, ls.lo);
  } else {
    if (fixed && a.small && !smooth_legacy()) {
      const int rows = std::max(4, (h + (int)grid.y - 1) / (int)grid.y);   // consecutive rows per block (2 extra per strip)
      const dim3 g2((w + kS2Out - 1) / kS2Out, (h + rows - 1) / rows, grid.z);
      if (pen == B2F_PENALTY_QUADRATIC) smooth2_lean_kernel<B2F_PENALTY_QUADRATIC><<<g2, kThreads, 0, st>>>(a, ls.lo, rows);
      else if (pen == B2F_PENALTY_L1) smooth2_lean_kernel<B2F_PENALTY_L1><<<g2, kThreads, 0, st>>>(a, ls.lo, rows);
      else smooth2_lean_kernel<B2F_PENALTY_LORENTZIAN><<<g2, kThreads, 0, st>>>(a, ls.lo, rows);
    } else if (fixed) {
      const int rows = std::max(4, (h + (int)grid.y - 1) / (int)grid.y);   // consecutive rows per block (2 extra per strip)
      const dim3 g2((w + kS2Out - 1) / kS2Out, (h + rows - 1) / rows, grid.z);
      if (pen == B2F_PENALTY_QUADRATIC) smooth2_kernel<B2F_PENALTY_QUADRATIC><<<g2, kThreads, 0, st>>>(a, ls.lo, rows);
      else if (pen == B2F_PENALTY_L1) smooth2_kernel<B2F_PENALTY_L1><<<g2, kThreads, 0, st>>>(a, ls.lo, rows);
      else smooth2_kernel<B2F_PENALTY_LORENTZIAN><<<g2, kThreads, 0, st>>>(a, ls.lo, rows);
    } else {
      if (pen == B2F_PENALTY_QUADRATIC) smooth2_generic_kernel<B2F_PENALTY_QUADRATIC, 0, 0><<<grid, kThreads, 0, st>>>(a, ls.lo);
      else if (pen == B2F_PENALTY_L1) smooth2_generic_kernel<B2F_PENALTY_L1, 0, 0><<<grid, kThreads, 0, st>>>(a, ls.lo);
      else smooth2_generic_kernel<B2F_PENALTY_LORENTZIAN, 0, 0><<<grid, kThreads, 0, st>>>(a, ls.lo);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) count_launch();
  rc = ls.end(loss_host);
  if (e != cudaSuccess) return cuda_fail(e, "smooth_kernel");
  return rc;
}

extern "C" int b2f_constvel_criterion(const float* f, const float* b, int B, int C, int h, int w,
                                      int size_average, float* grad_f, float* grad_b, double* loss_dev,
                                      double* loss_host, b2f_stream_t stream) {
  if (!f || !b) return fail(B2F_EINVAL, "constvel: NULL input");
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return fail(B2F_EINVAL, "constvel: bad size");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t hw = (int64_t)h * w, npix = (int64_t)B * hw;
  int blocks;
  dim3 grid;
  const bool vec = C == 2 && (hw & 3) == 0 && planes16(f) && planes16(b) && planes16(grad_f) && planes16(grad_b);
  int rc = grid_flat(B, vec ? hw / 4 : hw, &grid, &blocks);
  if (rc) return rc;
  // forward: 1/nElement (ConstVelCriterion.lua:33, 41-43); backward: 1/npixels (:58, 69-72)  (Q11)
  const double scale = size_average ? 1.0 / ((double)npix * C) : 1.0;
  const float gnorm = size_average ? (float)(1.0 / (double)npix) : 1.f;
  LossScratch ls;
  if ((rc = ls.begin(blocks, scale, loss_dev, st))) return rc;
  if (vec) constvel_c2_kernel<<<grid, kThreads, 0, st>>>(f, b, grad_f, grad_b, hw, gnorm, ls.lo);
  else constvel_kernel<<<grid, kThreads, 0, st>>>(f, b, grad_f, grad_b, B, C, hw, gnorm, ls.lo);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) count_launch();
  rc = ls.end(loss_host);
  if (e != cudaSuccess) return cuda_fail(e, "constvel_kernel");
  return rc;
}

extern "C" int b2f_occprior_criterion(const float* occ, int B, int C, int h, int w, float penalty,
                                      int size_average, float* grad, double* loss_dev, double* loss_host,
                                      b2f_stream_t stream) {
  if (!occ) return fail(B2F_EINVAL, "occprior: NULL input");
  if (B <= 0 || h <= 0 || w <= 0) return fail(B2F_EINVAL, "occprior: bad size");
  if (C != 2 && C != 3) return fail(B2F_EINVAL, "occprior: C=%d, expected 2 or 3", C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t hw = (int64_t)h * w, npix = (int64_t)B * hw;
  int blocks;
  dim3 grid;
  const bool vec = C == 2 && (hw & 3) == 0 && planes16(occ) && planes16(grad);
  int rc = grid_flat(B, vec ? hw / 4 : hw, &grid, &blocks);
  if (rc) return rc;
  const double scale = size_average ? 1.0 / (double)npix : 1.0;
  LossScratch ls;
  if ((rc = ls.begin(blocks, scale, loss_dev, st))) return rc;
  if (vec) occprior_c2_kernel<<<grid, kThreads, 0, st>>>(occ, grad, hw, penalty, (float)scale, ls.lo);
  else occprior_kernel<<<grid, kThreads, 0, st>>>(occ, grad, B, C, hw, penalty, (float)scale, ls.lo);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) count_launch();
  rc = ls.end(loss_host);
  if (e != cudaSuccess) return cuda_fail(e, "occprior_kernel");
  return rc;
}
