/* CPU restatement of the Back2Future hot path "as written"  --  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * PARITY UNPINNED: the reference (Lua/Torch7) cannot run in the build container and holds no
 * golden vectors for this path; see oracle/b2f_oracle.py for what pins the oracle instead.
 *
 * This file mirrors the reference's *algorithm structure* so that timing it is a fair stand-in
 * for the reference's own CPU path (Torch7 is not installable here):
 *   - cost volume: one pass per displacement -- elementwise multiply into a temporary, reduce
 *     over channels, accumulate into the window channel, final division
 *     (models/CostVolMulti.lua:62-100, 127-178);
 *   - sampler: per-pixel 4-tap gather / scatter with the CUDA kernel's pixel-offset + clamp
 *     semantics (extras/stnbhwd/BilinearSamplerBHWD.cu:6-20, 41-115, 161-307), NOT the
 *     normalised-coordinate maths of generic/BilinearSamplerBHWD.c.
 * OpenMP parallel loops stand in for TH's OpenMP tensor apply.  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

API int b2fcpu_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

API void b2fcpu_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Lua ranges of models/CostVolMulti.lua:77-88, converted to 0-based [r0, r0+len) for the ref
 * side and p0 for the frame side.  Returns 0 when the range is empty. */
static int win_range(int q, int size, int* r0, int* p0, int* len) {
  if (q < 0) { *r0 = 0; *p0 = -q; *len = size + q; }
  else       { *r0 = q; *p0 = 0;  *len = size - q; }
  return *len > 0;
}

/* CostVolMulti:updateOutput, models/CostVolMulti.lua:49-109 */
API int b2fcpu_costvol_forward(const float* const* frames, int F, int B, int C, int H, int W,
                               int win, int fwd, float* out) {
  const int n = (win - 1) / 2;
  const size_t hw = (size_t)H * W;
  memset(out, 0, sizeof(float) * (size_t)B * win * win * hw);
  float* tmp = (float*)malloc(sizeof(float) * (size_t)B * C * hw);   /* torch.cmul result */
  float* red = (float*)malloc(sizeof(float) * (size_t)B * hw);       /* cost:sum(2)       */
  if (!tmp || !red) { free(tmp); free(red); return -1; }
  const float* ref = frames[0];
  for (int f = 1; f < F; ++f) {
    const float* frame = frames[f];
    int i = 0;
    for (int qx_ = -n; qx_ <= n; ++qx_) {
      for (int qy_ = -n; qy_ <= n; ++qy_, ++i) {
        int qx = qx_ * f, qy = qy_ * f;
        if (!fwd) { qx = -qx; qy = -qy; }
        int rx, px, nx, ry, py, ny;
        if (!win_range(qx, W, &rx, &px, &nx) || !win_range(qy, H, &ry, &py, &ny)) continue;
        /* pass 1: cost = cmul(ref[qy,qx], frame[py,px]) */
#pragma omp parallel for collapse(2) schedule(static)
        for (int b = 0; b < B; ++b)
          for (int c = 0; c < C; ++c) {
            const float* r = ref + ((size_t)b * C + c) * hw;
            const float* g = frame + ((size_t)b * C + c) * hw;
            float* t = tmp + ((size_t)b * C + c) * (size_t)ny * nx;
            for (int y = 0; y < ny; ++y)
              for (int x = 0; x < nx; ++x)
                t[(size_t)y * nx + x] = r[(size_t)(ry + y) * W + rx + x] * g[(size_t)(py + y) * W + px + x];
          }
        /* pass 2: sum over channels */
#pragma omp parallel for schedule(static)
        for (int b = 0; b < B; ++b) {
          float* s = red + (size_t)b * ny * nx;
          for (size_t k = 0; k < (size_t)ny * nx; ++k) s[k] = 0.f;
          for (int c = 0; c < C; ++c) {
            const float* t = tmp + ((size_t)b * C + c) * (size_t)ny * nx;
            for (size_t k = 0; k < (size_t)ny * nx; ++k) s[k] += t[k];
          }
        }
        /* pass 3: output[{{},i,qy,qx}]:add(...) */
#pragma omp parallel for schedule(static)
        for (int b = 0; b < B; ++b) {
          const float* s = red + (size_t)b * ny * nx;
          float* o = out + ((size_t)b * win * win + i) * hw;
          for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
              o[(size_t)(ry + y) * W + rx + x] += s[(size_t)y * nx + x];
        }
      }
    }
  }
  const float k = (float)(C * (F - 1));
  const size_t total = (size_t)B * win * win * hw;
#pragma omp parallel for schedule(static)
  for (size_t j = 0; j < total; ++j) out[j] /= k;
  free(tmp); free(red);
  return 0;
}

/* CostVolMulti:updateGradInput, models/CostVolMulti.lua:111-181.
 * gradOut may be a narrow of a wider buffer: batch stride in elements (0 = contiguous). */
API int b2fcpu_costvol_backward(const float* const* frames, int F, int B, int C, int H, int W,
                                int win, int fwd, const float* gradOut, int64_t go_bstride,
                                float* const* gradFrames) {
  const int n = (win - 1) / 2;
  const size_t hw = (size_t)H * W;
  if (go_bstride == 0) go_bstride = (int64_t)win * win * hw;
  for (int f = 0; f < F; ++f) memset(gradFrames[f], 0, sizeof(float) * (size_t)B * C * hw);
  const float* ref = frames[0];
  float* gref = gradFrames[0];
  for (int f = 1; f < F; ++f) {
    const float* frame = frames[f];
    float* gfr = gradFrames[f];
    int i = 0;
    for (int qx_ = -n; qx_ <= n; ++qx_) {
      for (int qy_ = -n; qy_ <= n; ++qy_, ++i) {
        int qx = qx_ * f, qy = qy_ * f;
        if (!fwd) { qx = -qx; qy = -qy; }
        int rx, px, nx, ry, py, ny;
        if (!win_range(qx, W, &rx, &px, &nx) || !win_range(qy, H, &ry, &py, &ny)) continue;
#pragma omp parallel for collapse(2) schedule(static)
        for (int b = 0; b < B; ++b)
          for (int c = 0; c < C; ++c) {
            const float* go = gradOut + (size_t)b * go_bstride + (size_t)i * hw;
            const float* r = ref + ((size_t)b * C + c) * hw;
            const float* g = frame + ((size_t)b * C + c) * hw;
            float* gr = gref + ((size_t)b * C + c) * hw;
            float* gf = gfr + ((size_t)b * C + c) * hw;
            for (int y = 0; y < ny; ++y)
              for (int x = 0; x < nx; ++x) {
                const float v = go[(size_t)(ry + y) * W + rx + x];
                gr[(size_t)(ry + y) * W + rx + x] += v * g[(size_t)(py + y) * W + px + x];
                gf[(size_t)(py + y) * W + px + x] += v * r[(size_t)(ry + y) * W + rx + x];
              }
          }
      }
    }
  }
  const float k = (float)(C * (F - 1));
  for (int f = 0; f < F; ++f) {
    float* g = gradFrames[f];
    const size_t total = (size_t)B * C * hw;
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < total; ++j) g[j] /= k;
  }
  return 0;
}

/* getTopLeft, BilinearSamplerBHWD.cu:6-20 */
static inline void top_left(float off, int idx, int size, int* point, float* weight) {
  float xc = off + (float)idx;
  if (xc < 0.f) xc = 0.f;
  if (xc > (float)(size - 1)) xc = (float)(size - 1);
  const float fl = floorf(xc);
  *point = (int)fl;
  *weight = 1.f - (xc - fl);
}

/* bilinearSamplingFromGrid, BilinearSamplerBHWD.cu:41-115 */
API int b2fcpu_warp_forward(const float* img, const float* grid, float* out,
                            int B, int H, int W, int C, int Hg, int Wg) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int yo = 0; yo < Hg; ++yo)
      for (int xo = 0; xo < Wg; ++xo) {
        const float* g = grid + (((size_t)b * Hg + yo) * Wg + xo) * 2;
        int xi, yi; float wx, wy;
        top_left(g[0], xo, W, &xi, &wx);
        top_left(g[1], yo, H, &yi, &wy);
        const int rin = xi + 1 <= W - 1, bin = yi + 1 <= H - 1;
        const float* tl = img + (((size_t)b * H + yi) * W + xi) * C;
        const float* tr = tl + C;
        const float* bl = tl + (size_t)W * C;
        const float* br = bl + C;
        float* o = out + (((size_t)b * Hg + yo) * Wg + xo) * C;
        for (int c = 0; c < C; ++c) {
          const float vtl = tl[c];
          const float vtr = rin ? tr[c] : 0.f;
          const float vbl = bin ? bl[c] : 0.f;
          const float vbr = (rin && bin) ? br[c] : 0.f;
          o[c] = wx * wy * vtl + (1 - wx) * wy * vtr + wx * (1 - wy) * vbl + (1 - wx) * (1 - wy) * vbr;
        }
      }
  return 0;
}

/* backwardBilinearSampling<onlyGrid>, BilinearSamplerBHWD.cu:161-307.  gradImg == NULL selects
 * onlyGrid.  gradImg is ACCUMULATED into (the Lua wrapper zero-fills it first, :99-102); the
 * scatter is serialised per batch element so that no atomics are needed on the host. */
API int b2fcpu_warp_backward(const float* img, const float* grid, const float* gradOut,
                             float* gradImg, float* gradGrid,
                             int B, int H, int W, int C, int Hg, int Wg) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b)
    for (int yo = 0; yo < Hg; ++yo)
      for (int xo = 0; xo < Wg; ++xo) {
        const size_t gi = (((size_t)b * Hg + yo) * Wg + xo);
        const float* g = grid + gi * 2;
        int xi, yi; float wx, wy;
        top_left(g[0], xo, W, &xi, &wx);
        top_left(g[1], yo, H, &yi, &wy);
        const int rin = xi + 1 <= W - 1, bin = yi + 1 <= H - 1;
        const size_t a = (((size_t)b * H + yi) * W + xi) * C;
        const float* go = gradOut + gi * C;
        float dtl = 0, dtr = 0, dbl = 0, dbr = 0;
        for (int c = 0; c < C; ++c) {
          const float v = go[c];
          dtl += img[a + c] * v;
          if (gradImg) gradImg[a + c] += wx * wy * v;
          if (rin) {
            dtr += img[a + C + c] * v;
            if (gradImg) gradImg[a + C + c] += (1 - wx) * wy * v;
          }
          if (bin) {
            dbl += img[a + (size_t)W * C + c] * v;
            if (gradImg) gradImg[a + (size_t)W * C + c] += wx * (1 - wy) * v;
          }
          if (rin && bin) {
            dbr += img[a + (size_t)W * C + C + c] * v;
            if (gradImg) gradImg[a + (size_t)W * C + C + c] += (1 - wx) * (1 - wy) * v;
          }
        }
        gradGrid[gi * 2 + 0] = -wy * dtl + wy * dtr - (1 - wy) * dbl + (1 - wy) * dbr;
        gradGrid[gi * 2 + 1] = -wx * dtl + wx * dbl - (1 - wx) * dtr + (1 - wx) * dbr;
      }
  return 0;
}
