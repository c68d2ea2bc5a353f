/* Double-precision closed-form checker of the cost volume and the sampler  --  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * PARITY UNPINNED for the cost volume (see oracle/b2f_oracle.py); the sampler maths is the one pinned against the
 * reference's own kernel in tests/test_ref_sampler.py.
 *
 * oracle/b2f_oracle.py (numpy, float64) is the oracle; it takes tens of seconds at the benchmark's full sizes
 * (B = 8, 1024x448 pyramid).  This file restates the same closed forms with double accumulation and OpenMP so that
 * EVERY op of the benchmarked step can be compared at full size in seconds (tests/test_bench_parity.py and the
 * `parity` block of bench.py); tests/test_oracle.py pins it to the numpy oracle on small ragged cases.
 *   cost volume: models/CostVolMulti.lua:49-109 (forward), :111-181 (backward); SURVEY Q4 conventions
 *   sampler:     extras/stnbhwd/BilinearSamplerBHWD.cu:6-20 (geometry, fp32 exactly as the kernel), :41-115, :161-307
 * Inputs are float32 (what the kernels see); every product and sum is in double.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* out[b, i, y, x] = 1/(C (F-1)) sum_f sum_c ref[b,c,y,x] * frame_f[b,c,y - s f qy, x - s f qx], i = (qx+n) win + (qy+n),
 * s = +1 (fwd) / -1, out-of-range sources dropped (CostVolMulti.lua:65-100). */
API int b2fchk_costvol_forward(const float* const* frames, int F, int B, int C, int H, int W,
                               int win, int fwd, double* out) {
  const int n = (win - 1) / 2, s = fwd ? 1 : -1;
  const size_t hw = (size_t)H * W;
  const double k = 1.0 / ((double)C * (F - 1));
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < win * win; ++i) {
      const int qx_ = i / win - n, qy_ = i % win - n;
      double* o = out + ((size_t)b * win * win + i) * hw;
      for (size_t j = 0; j < hw; ++j) o[j] = 0.0;
      for (int f = 1; f < F; ++f) {
        const int qx = s * f * qx_, qy = s * f * qy_;
        for (int c = 0; c < C; ++c) {
          const float* r = frames[0] + ((size_t)b * C + c) * hw;
          const float* g = frames[f] + ((size_t)b * C + c) * hw;
          for (int y = 0; y < H; ++y) {
            const int ys = y - qy;
            if (ys < 0 || ys >= H) continue;
            for (int x = 0; x < W; ++x) {
              const int xs = x - qx;
              if (xs < 0 || xs >= W) continue;
              o[(size_t)y * W + x] += (double)r[(size_t)y * W + x] * (double)g[(size_t)ys * W + xs];
            }
          }
        }
      }
      for (size_t j = 0; j < hw; ++j) o[j] *= k;
    }
  return 0;
}

/* gradRef[b,c,p] = k sum_f sum_i go[b,i,p] frame_f[b,c,p - q];  gradFrame_f[b,c,p'] = k sum_i go[b,i,p'+q] ref[b,c,p'+q]
 * (CostVolMulti.lua:127-178).  gradFrames[f] may be NULL. */
API int b2fchk_costvol_backward(const float* const* frames, int F, int B, int C, int H, int W,
                                int win, int fwd, const float* gradOut, int64_t go_bstride,
                                double* const* gradFrames) {
  const int n = (win - 1) / 2, s = fwd ? 1 : -1;
  const size_t hw = (size_t)H * W;
  const double k = 1.0 / ((double)C * (F - 1));
  if (go_bstride == 0) go_bstride = (int64_t)win * win * hw;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* r = frames[0] + ((size_t)b * C + c) * hw;
      double* gr = gradFrames[0] ? gradFrames[0] + ((size_t)b * C + c) * hw : NULL;
      if (gr) for (size_t j = 0; j < hw; ++j) gr[j] = 0.0;
      for (int f = 1; f < F; ++f) {
        const float* g = frames[f] + ((size_t)b * C + c) * hw;
        double* gf = gradFrames[f] ? gradFrames[f] + ((size_t)b * C + c) * hw : NULL;
        if (gf) for (size_t j = 0; j < hw; ++j) gf[j] = 0.0;
        for (int i = 0; i < win * win; ++i) {
          const int qx = s * f * (i / win - n), qy = s * f * (i % win - n);
          const float* go = gradOut + (size_t)b * go_bstride + (size_t)i * hw;
          for (int y = 0; y < H; ++y) {
            const int ys = y - qy;
            if (ys < 0 || ys >= H) continue;
            for (int x = 0; x < W; ++x) {
              const int xs = x - qx;
              if (xs < 0 || xs >= W) continue;
              const double v = (double)go[(size_t)y * W + x];
              if (gr) gr[(size_t)y * W + x] += v * (double)g[(size_t)ys * W + xs];
              if (gf) gf[(size_t)ys * W + xs] += v * (double)r[(size_t)y * W + x];
            }
          }
        }
        if (gf) for (size_t j = 0; j < hw; ++j) gf[j] *= k;
      }
      if (gr) for (size_t j = 0; j < hw; ++j) gr[j] *= k;
    }
  return 0;
}

/* getTopLeft, BilinearSamplerBHWD.cu:6-20: three separately rounded fp32 operations, then floor */
static inline void top_left(float off, int idx, int size, int* point, double* weight) {
  volatile float xc = off + (float)idx;
  if (xc < 0.f) xc = 0.f;
  if (xc > (float)(size - 1)) xc = (float)(size - 1);
  const float fl = floorf(xc);
  volatile float frac = xc - fl;
  volatile float w = 1.f - frac;
  *point = (int)fl;
  *weight = (double)w;
}

/* bilinearSamplingFromGrid, BilinearSamplerBHWD.cu:41-115 */
API int b2fchk_warp_forward(const float* img, const float* grid, double* out,
                            int B, int H, int W, int C, int Hg, int Wg) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int yo = 0; yo < Hg; ++yo)
      for (int xo = 0; xo < Wg; ++xo) {
        const float* g = grid + (((size_t)b * Hg + yo) * Wg + xo) * 2;
        int xi, yi;
        double wx, wy;
        top_left(g[0], xo, W, &xi, &wx);
        top_left(g[1], yo, H, &yi, &wy);
        const int rin = xi + 1 <= W - 1, bin = yi + 1 <= H - 1;
        const float* tl = img + (((size_t)b * H + yi) * W + xi) * C;
        double* o = out + (((size_t)b * Hg + yo) * Wg + xo) * C;
        for (int c = 0; c < C; ++c) {
          const double vtl = tl[c];
          const double vtr = rin ? tl[C + c] : 0.0;
          const double vbl = bin ? tl[(size_t)W * C + c] : 0.0;
          const double vbr = (rin && bin) ? tl[(size_t)W * C + C + c] : 0.0;
          o[c] = wx * wy * vtl + (1 - wx) * wy * vtr + wx * (1 - wy) * vbl + (1 - wx) * (1 - wy) * vbr;
        }
      }
  return 0;
}

/* backwardBilinearSampling<onlyGrid>, BilinearSamplerBHWD.cu:161-307; gradImg (zero-filled here, then accumulated)
 * may be NULL (= onlyGrid).  The scatter is serial per batch item. */
API int b2fchk_warp_backward(const float* img, const float* grid, const float* gradOut,
                             double* gradImg, double* gradGrid,
                             int B, int H, int W, int C, int Hg, int Wg) {
  if (gradImg) memset(gradImg, 0, sizeof(double) * (size_t)B * H * W * C);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b)
    for (int yo = 0; yo < Hg; ++yo)
      for (int xo = 0; xo < Wg; ++xo) {
        const size_t gi = (((size_t)b * Hg + yo) * Wg + xo);
        const float* g = grid + gi * 2;
        int xi, yi;
        double wx, wy;
        top_left(g[0], xo, W, &xi, &wx);
        top_left(g[1], yo, H, &yi, &wy);
        const int rin = xi + 1 <= W - 1, bin = yi + 1 <= H - 1;
        const size_t a = (((size_t)b * H + yi) * W + xi) * C;
        const size_t dn = (size_t)W * C;
        const float* go = gradOut + gi * C;
        double dtl = 0, dtr = 0, dbl = 0, dbr = 0;
        for (int c = 0; c < C; ++c) {
          const double v = go[c];
          dtl += (double)img[a + c] * v;
          if (gradImg) gradImg[a + c] += wx * wy * v;
          if (rin) {
            dtr += (double)img[a + C + c] * v;
            if (gradImg) gradImg[a + C + c] += (1 - wx) * wy * v;
          }
          if (bin) {
            dbl += (double)img[a + dn + c] * v;
            if (gradImg) gradImg[a + dn + c] += wx * (1 - wy) * v;
          }
          if (rin && bin) {
            dbr += (double)img[a + dn + C + c] * v;
            if (gradImg) gradImg[a + dn + C + c] += (1 - wx) * (1 - wy) * v;
          }
        }
        gradGrid[gi * 2 + 0] = -wy * dtl + wy * dtr - (1 - wy) * dbl + (1 - wy) * dbr;
        gradGrid[gi * 2 + 1] = -wx * dtl + wx * dbl - (1 - wx) * dtr + (1 - wx) * dbr;
      }
  return 0;
}

/* max |a - b| / max(|b|, rms(b)) (SURVEY 8c tolerance), a float32 (kernel result, element stride 1 inside rows of
 * `row` elements, row stride `a_rstride`), b float64 dense.  Returns the maximum; *worst receives its flat index. */
API double b2fchk_rel_err(const float* a, int64_t rows, int64_t row, int64_t a_rstride, const double* b, int64_t* worst) {
  const int64_t total = rows * row;
  double ss = 0.0;
#pragma omp parallel for reduction(+ : ss) schedule(static)
  for (int64_t j = 0; j < total; ++j) ss += b[j] * b[j];
  const double rms = sqrt(ss / (double)(total > 0 ? total : 1)) + 1e-30;
  double best = 0.0;
  int64_t at = -1;
#pragma omp parallel
  {
    double lb = 0.0;
    int64_t la = -1;
#pragma omp for schedule(static) nowait
    for (int64_t r = 0; r < rows; ++r)
      for (int64_t x = 0; x < row; ++x) {
        const double bv = b[r * row + x];
        const double av = (double)a[r * a_rstride + x];
        double sc = fabs(bv);
        if (sc < rms) sc = rms;
        double e = fabs(av - bv) / sc;
        if (!(e == e)) e = INFINITY;   /* NaN in the kernel result */
        if (e > lb) { lb = e; la = r * row + x; }
      }
#pragma omp critical
    if (lb > best) { best = lb; at = la; }
  }
  if (worst) *worst = at;
  return best;
}
