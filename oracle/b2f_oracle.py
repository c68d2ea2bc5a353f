"""CPU oracle for the Back2Future hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This module is a numpy restatement of the reference's behaviour (quirks included) for the
path BASELINE.json's north_star names.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
(``back2future_b200``) never does and fails loudly when its CUDA library is missing.

PARITY: the sampler (warp_forward / warp_backward, rows a4/a5) is PINNED to the reference itself --
the reference's own CUDA sampler, compiled unmodified into oracle/_ref/libstn_ref.so, produced
tests/golden/ref_sampler_golden.npz on a B200 and this module reproduces it to 1e-6
(tests/test_golden.py, tests/test_ref_sampler.py).  Everything else here (cost volume, criterions:
Lua over Torch7 tensor methods) is PARITY UNPINNED: the reference cannot be executed in the build
container (no LuaJIT, no Torch7, no network) and its own tests hold no golden vector, known-answer
value or fixture for those rows.  What pins those parts instead:
  * two independent restatements per module where that is possible (a literal transcription
    of the Lua loops next to a closed form) cross-checked in tests/test_oracle.py;
  * finite-difference gradient checks in the spirit of the reference's commented-out
    Jacobian tests (models/CostVolMulti.lua:192-223, extras/stnbhwd/test.lua:47-120);
  * derived known-answer cases (delta images from models/CostVolMulti.lua:225-254, integer
    shifts and identity flow for the sampler);
  * an independent C restatement (oracle/c/b2f_cpu.c) compared element-wise.

All citations are relative to /root/reference.

Conventions
-----------
* Arrays are numpy, layouts as in the reference: feature maps / criterion inputs are BDHW
  (B, C, h, w); sampler images are BHWD (B, H, W, C); sampler grids are (B, Hg, Wg, 2) with
  channel 0 = x offset and channel 1 = y offset in PIXELS (BilinearSamplerBHWD.cu:69-70).
* ``dtype`` selects the arithmetic type of the restatement: float64 gives the mathematically
  exact value of the reference's formula on the given fp32 inputs (used with the 1e-4
  relative tolerance of north_star), float32 rounds after every tensor operation the way the
  chain of Torch7 kernels does.
* Decisions that are DIS-continuous in the inputs (floor / clamp in the sampler, the
  out-of-image mask of OBCC/OBGCC) are always evaluated in float32 with the reference's
  operation order, whatever ``dtype`` is, because there the rounding IS the behaviour.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# ----------------------------------------------------------------------------------------
# penalties  (criterions/penalty/*.lua)
# ----------------------------------------------------------------------------------------

PEN_QUADRATIC = 0
PEN_L1 = 1
PEN_LORENTZIAN = 2


class QuadraticPenalty:
    """criterions/penalty/quadratic_function.lua:15-21."""
    kind = PEN_QUADRATIC
    eps = 0.0

    def apply(self, x):
        return x * x

    def der(self, x):
        return x * x.dtype.type(2)


class L1Penalty:
    """criterions/penalty/L1_function.lua:15-26.

    ``self.alpha = 0.5 or alpha`` always evaluates to 0.5 in Lua, so the constructor
    argument is ignored: p(x) = sqrt(x^2 + 1e-6), p'(x) = x / sqrt(x^2 + 1e-6).
    """
    kind = PEN_L1

    def __init__(self, alpha=None):
        self.eps = 0.001 * 0.001
        self.alpha = 0.5

    def apply(self, x):
        t = x.dtype.type
        return np.power(x * x + t(self.eps), t(self.alpha))

    def der(self, x):
        t = x.dtype.type
        return (x * t(2 * self.alpha)) / np.power(x * x + t(self.eps), t(1 - self.alpha))


class LorentzianPenalty:
    """criterions/penalty/Lorentzian_function.lua:15-31: log(1 + x^2/(2 eps^2)), 2x/(x^2+2eps^2)."""
    kind = PEN_LORENTZIAN

    def __init__(self):
        self.set_eps(0.05)

    def set_eps(self, eps):
        self.eps = eps
        self.eps_sq = eps * eps

    def apply(self, x):
        t = x.dtype.type
        return np.log(t(1) + t(0.5) * ((x * x) / t(self.eps_sq)))

    def der(self, x):
        t = x.dtype.type
        return (x * t(2)) / (x * x + t(2 * self.eps_sq))


def make_penalty(kind, eps=None):
    if kind == PEN_QUADRATIC:
        return QuadraticPenalty()
    if kind == PEN_L1:
        return L1Penalty()
    if kind == PEN_LORENTZIAN:
        p = LorentzianPenalty()
        if eps is not None:
            p.set_eps(eps)
        return p
    raise ValueError("unknown penalty kind %r" % (kind,))


# ----------------------------------------------------------------------------------------
# nn.CostVolMulti  (models/CostVolMulti.lua)
# ----------------------------------------------------------------------------------------

def _lua_ranges(q, size):
    """The (ref range, frame range) pair of models/CostVolMulti.lua:77-88, 1-based inclusive.

    Returns python slices, or None when Torch7 would raise on an inverted range
    (|q| >= size; never happens at the model's shapes)."""
    if q < 0:
        qr, pr = (1, size + q), (1 - q, size)
    else:
        qr, pr = (1 + q, size), (1, size - q)
    if qr[0] > qr[1]:
        return None
    return slice(qr[0] - 1, qr[1]), slice(pr[0] - 1, pr[1])


def costvol_forward_lua(frames, win=9, fwd=True, dtype=np.float64):
    """Literal transcription of CostVolMulti:updateOutput (models/CostVolMulti.lua:49-109).

    Loop order, channel index ``i`` (x-major: q_x outer, q_y inner, :66-67), per-frame
    displacement multiplier (f-1) (:68-69), window mirroring for ``fwd == false`` (:71-74)
    and the constant normaliser N*(frames-1) (:100) are kept exactly."""
    frames = [np.asarray(f, dtype=dtype) for f in frames]
    nfr = len(frames)
    ref = frames[0]
    B, N, h, w = ref.shape
    n = (win - 1) // 2
    out = np.zeros((B, win * win, h, w), dtype=dtype)
    for f in range(2, nfr + 1):
        frame = frames[f - 1]
        i = 0
        for q_x_ in range(-n, n + 1):
            for q_y_ in range(-n, n + 1):
                q_x = q_x_ * (f - 1)
                q_y = q_y_ * (f - 1)
                if not fwd:
                    q_x, q_y = -q_x, -q_y
                rx = _lua_ranges(q_x, w)
                ry = _lua_ranges(q_y, h)
                if rx is not None and ry is not None:
                    (qx, px), (qy, py) = rx, ry
                    cost = ref[:, :, qy, qx] * frame[:, :, py, px]
                    out[:, i, qy, qx] += cost.sum(axis=1, dtype=dtype)
                i += 1
    out /= dtype(N * (nfr - 1))
    return out


def costvol_forward(frames, win=9, fwd=True, dtype=np.float64):
    """Closed form of the same module:

        out[b, (qx+n)*win + (qy+n), y, x] =
            1/(C (F-1)) * sum_f sum_c ref[b,c,y,x] * frame_f[b,c, y - s m qy, x - s m qx]

    with s = +1 (fwd) / -1, m = f-1 and out-of-range sources dropped.  Written with an
    explicit zero-padded copy of the frame so that it shares no indexing code with
    ``costvol_forward_lua``."""
    frames = [np.asarray(f, dtype=dtype) for f in frames]
    nfr = len(frames)
    ref = frames[0]
    B, C, h, w = ref.shape
    n = (win - 1) // 2
    s = 1 if fwd else -1
    out = np.zeros((B, win * win, h, w), dtype=dtype)
    for m in range(1, nfr):
        pad = n * m
        fp = np.zeros((B, C, h + 2 * pad, w + 2 * pad), dtype=dtype)
        fp[:, :, pad:pad + h, pad:pad + w] = frames[m]
        for ix in range(win):
            for iy in range(win):
                dx = s * m * (ix - n)
                dy = s * m * (iy - n)
                src = fp[:, :, pad - dy:pad - dy + h, pad - dx:pad - dx + w]
                out[:, ix * win + iy] += np.einsum('bchw,bchw->bhw', ref, src)
    out /= dtype(C * (nfr - 1))
    return out


def costvol_backward(frames, grad_out, win=9, fwd=True, dtype=np.float64):
    """CostVolMulti:updateGradInput (models/CostVolMulti.lua:111-181), closed form.

        gradRef[b,c,y,x]      = k sum_f sum_i go[b,i,y,x]          * frame_f[b,c,y-dy,x-dx]
        gradFrame_f[b,c,y',x'] = k       sum_i go[b,i,y'+dy,x'+dx] * ref[b,c,y'+dy,x'+dx]

    k = 1/(C (F-1)).  Returns a list of F arrays like ``frames``."""
    frames = [np.asarray(f, dtype=dtype) for f in frames]
    go = np.asarray(grad_out, dtype=dtype)
    nfr = len(frames)
    ref = frames[0]
    B, C, h, w = ref.shape
    n = (win - 1) // 2
    s = 1 if fwd else -1
    grads = [np.zeros_like(ref) for _ in range(nfr)]
    for m in range(1, nfr):
        frame = frames[m]
        for ix in range(win):
            for iy in range(win):
                dx = s * m * (ix - n)
                dy = s * m * (iy - n)
                # destination (ref) window and source (frame) window, 0-based
                y0, y1 = max(0, dy), min(h, h + dy)
                x0, x1 = max(0, dx), min(w, w + dx)
                if y0 >= y1 or x0 >= x1:
                    continue
                g = go[:, ix * win + iy, y0:y1, x0:x1][:, None]
                grads[0][:, :, y0:y1, x0:x1] += g * frame[:, :, y0 - dy:y1 - dy, x0 - dx:x1 - dx]
                grads[m][:, :, y0 - dy:y1 - dy, x0 - dx:x1 - dx] += g * ref[:, :, y0:y1, x0:x1]
    k = dtype(C * (nfr - 1))
    return [g / k for g in grads]


# ----------------------------------------------------------------------------------------
# nn.BilinearSamplerBHWD  (extras/stnbhwd/BilinearSamplerBHWD.cu -- the CUDA semantics, Q1)
# ----------------------------------------------------------------------------------------

def _top_left(offset, idx, size):
    """getTopLeft (BilinearSamplerBHWD.cu:6-20), always in float32.

    xcoord = offset + idx ; clamp to [0, size-1] ; point = floor ; weight = 1-(xcoord-point).
    ``offset`` (..., ) float32 array, ``idx`` broadcastable int array."""
    xc = (np.asarray(offset, F32) + np.asarray(idx).astype(F32)).astype(F32)
    xc = np.where(xc < F32(0), F32(0), xc)
    xc = np.where(xc > F32(size - 1), F32(size - 1), xc).astype(F32)
    pt = np.floor(xc)
    wgt = (F32(1) - (xc - pt)).astype(F32)
    return pt.astype(np.int64), wgt


def _warp_geometry(img_shape, grid):
    B, H, W, C = img_shape
    _, Hg, Wg, _ = grid.shape
    xo = np.arange(Wg)[None, None, :]
    yo = np.arange(Hg)[None, :, None]
    xi, wx = _top_left(grid[..., 0], xo, W)
    yi, wy = _top_left(grid[..., 1], yo, H)
    # BilinearSamplerBHWD.cu:89-92: a tap at index W (or H) is "not in" and contributes 0
    right_in = (xi + 1) <= (W - 1)
    bottom_in = (yi + 1) <= (H - 1)
    return xi, wx, yi, wy, right_in, bottom_in


def warp_forward(img, grid, dtype=np.float64):
    """bilinearSamplingFromGrid (BilinearSamplerBHWD.cu:41-115) + Lua wrapper (:53-79).

    img (B,H,W,C), grid (B,Hg,Wg,2) pixel offsets (x, y); returns (B,Hg,Wg,C)."""
    img32 = np.asarray(img, F32)
    grid = np.asarray(grid, F32)
    B, H, W, C = img32.shape
    xi, wx, yi, wy, rin, bin_ = _warp_geometry(img32.shape, grid)
    im = img32.astype(dtype)
    bidx = np.arange(B)[:, None, None]
    x1 = np.minimum(xi + 1, W - 1)
    y1 = np.minimum(yi + 1, H - 1)
    tl = im[bidx, yi, xi]
    tr = im[bidx, yi, x1] * rin[..., None]
    bl = im[bidx, y1, xi] * bin_[..., None]
    br = im[bidx, y1, x1] * (rin & bin_)[..., None]
    wx = wx.astype(dtype)[..., None]
    wy = wy.astype(dtype)[..., None]
    one = dtype(1)
    return (wx * wy * tl + (one - wx) * wy * tr + wx * (one - wy) * bl
            + (one - wx) * (one - wy) * br).astype(dtype)


def warp_backward(img, grid, grad_out, only_grid=False, dtype=np.float64):
    """backwardBilinearSampling<onlyGrid> (BilinearSamplerBHWD.cu:161-307).

    Returns (gradImg or None, gradGrid).  gradImg is the scatter-add of the four tap weights
    times gradOut into in-bounds taps (:236-262); gradGrid = (x, y) built from the four
    per-pixel dot products (:287-295) with no clamp derivative and no (W-1)/2 scaling."""
    img32 = np.asarray(img, F32)
    grid = np.asarray(grid, F32)
    go = np.asarray(grad_out).astype(dtype)
    B, H, W, C = img32.shape
    xi, wx, yi, wy, rin, bin_ = _warp_geometry(img32.shape, grid)
    im = img32.astype(dtype)
    bidx = np.broadcast_to(np.arange(B)[:, None, None], xi.shape)
    x1 = np.minimum(xi + 1, W - 1)
    y1 = np.minimum(yi + 1, H - 1)
    wxd = wx.astype(dtype)
    wyd = wy.astype(dtype)
    one = dtype(1)
    taps = (
        (yi, xi, np.ones_like(rin), wxd * wyd),
        (yi, x1, rin, (one - wxd) * wyd),
        (y1, xi, bin_, wxd * (one - wyd)),
        (y1, x1, rin & bin_, (one - wxd) * (one - wyd)),
    )
    grad_img = None
    if not only_grid:
        grad_img = np.zeros(img32.shape, dtype=dtype)
    dots = []
    for (ty, tx, tin, tw) in taps:
        v = im[bidx, ty, tx] * tin[..., None]
        dots.append((v * go).sum(axis=-1))
        if not only_grid:
            np.add.at(grad_img, (bidx, ty, tx), (tw * tin)[..., None] * go)
    d_tl, d_tr, d_bl, d_br = dots
    gx = -wyd * d_tl + wyd * d_tr - (one - wyd) * d_bl + (one - wyd) * d_br
    gy = -wxd * d_tl + wxd * d_bl - (one - wxd) * d_tr + (one - wxd) * d_br
    grad_grid = np.stack([gx, gy], axis=-1).astype(dtype)
    return grad_img, grad_grid


def warping_unit_forward(img_bdhw, flow_bdhw, scale, dtype=np.float64):
    """models/pwc.lua:68-73 `warpingUnit(I, F)` behind the `nn.MulConstant(scale)` of :402-408 / :441-446:
    Transpose({2,3},{3,4}) of image and flow (BDHW -> BHWD), BilinearSamplerBHWD, Transpose({3,4},{2,3}) back.
    img (B,C,H,W), flow (B,2,H,W) in network units; returns (B,C,H,W).  The product flow*scale is rounded to
    fp32 (MulConstant runs in fp32) before the sampler sees it."""
    img = np.asarray(img_bdhw, F32)
    grid = (np.asarray(flow_bdhw, F32) * F32(scale)).astype(F32)
    out = warp_forward(np.ascontiguousarray(img.transpose(0, 2, 3, 1)),
                       np.ascontiguousarray(grid.transpose(0, 2, 3, 1)), dtype=dtype)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2))


def warping_unit_backward(img_bdhw, flow_bdhw, scale, grad_out_bdhw, only_grid=False, dtype=np.float64):
    """Backward of the same chain: (gradImg (B,C,H,W) or None, gradFlow (B,2,H,W)); MulConstant's backward
    multiplies the sampler's flow gradient by `scale` (nn.MulConstant.updateGradInput)."""
    img = np.asarray(img_bdhw, F32)
    grid = (np.asarray(flow_bdhw, F32) * F32(scale)).astype(F32)
    go = np.ascontiguousarray(np.asarray(grad_out_bdhw).transpose(0, 2, 3, 1))
    gi, gg = warp_backward(np.ascontiguousarray(img.transpose(0, 2, 3, 1)),
                           np.ascontiguousarray(grid.transpose(0, 2, 3, 1)), go, only_grid=only_grid, dtype=dtype)
    gflow = np.ascontiguousarray(gg.transpose(0, 3, 1, 2)) * dtype(F32(scale))
    return (None if gi is None else np.ascontiguousarray(gi.transpose(0, 3, 1, 2))), gflow


def warp_forward_loops(img, grid):
    """Scalar per-pixel transcription of the kernel (float64 arithmetic, fp32 geometry).
    Slow: small cases only.  Independent of the vectorised version above."""
    img = np.asarray(img, F32)
    grid = np.asarray(grid, F32)
    B, H, W, C = img.shape
    _, Hg, Wg, _ = grid.shape
    out = np.zeros((B, Hg, Wg, C))
    for b in range(B):
        for yo in range(Hg):
            for xo in range(Wg):
                xc = F32(grid[b, yo, xo, 0] + F32(xo))
                yc = F32(grid[b, yo, xo, 1] + F32(yo))
                xc = F32(min(max(xc, F32(0)), F32(W - 1)))
                yc = F32(min(max(yc, F32(0)), F32(H - 1)))
                xi, yi = int(np.floor(xc)), int(np.floor(yc))
                wx = float(F32(1) - F32(xc - F32(xi)))
                wy = float(F32(1) - F32(yc - F32(yi)))
                for c in range(C):
                    tl = float(img[b, yi, xi, c])
                    tr = float(img[b, yi, xi + 1, c]) if xi + 1 <= W - 1 else 0.0
                    bl = float(img[b, yi + 1, xi, c]) if yi + 1 <= H - 1 else 0.0
                    br = float(img[b, yi + 1, xi + 1, c]) if (xi + 1 <= W - 1 and yi + 1 <= H - 1) else 0.0
                    out[b, yo, xo, c] = (wx * wy * tl + (1 - wx) * wy * tr
                                         + wx * (1 - wy) * bl + (1 - wx) * (1 - wy) * br)
    return out


# ----------------------------------------------------------------------------------------
# OBCC / OBGCC  (criterions/OBCCriterion.lua, criterions/OBGCCriterion.lua)
# ----------------------------------------------------------------------------------------

def _frame_roles(F, past_flow):
    """For warped frame f = 1..F-1: (occlusion channel 0-based, flow input index 0-based,
    multiplier k).  OBCCriterion.lua:79-93: past frames (f <= ref) use occ[:,2], flow
    input[2] if past_flow else input[1], k = f-ref-1; future frames use occ[:,1], input[1],
    k = f-ref.  ref = 0.5 (F-1)."""
    ref = 0.5 * (F - 1)
    roles = []
    for f in range(1, F):
        if f <= ref:
            roles.append((1, 1 if past_flow else 0, f - ref - 1))
        else:
            roles.append((0, 0, f - ref))
    return roles


def out_of_image_mask(flow, k, scale):
    """OBCCriterion.lua:54-56, 81-89, 97-100 -- always float32, three separately rounded ops
    (SURVEY Q14): tcoord = fl(coord + fl(fl(k*flow)*scale)); mask = 1<=x<=w and 1<=y<=h with
    1-based coordinates.  flow (B,2,h,w).  Returns float mask (B,h,w) in {0,1}."""
    flow = np.asarray(flow, F32)
    B, _, h, w = flow.shape
    cx = np.arange(1, w + 1, dtype=F32)[None, None, :]
    cy = np.arange(1, h + 1, dtype=F32)[None, :, None]
    kf = F32(k)
    sc = F32(scale)
    tx = (cx + ((flow[:, 0] * kf).astype(F32) * sc).astype(F32)).astype(F32)
    ty = (cy + ((flow[:, 1] * kf).astype(F32) * sc).astype(F32)).astype(F32)
    m = (tx >= F32(1)) & (ty >= F32(1)) & (tx <= F32(w)) & (ty <= F32(h))
    return m


def _fwd_diff(a, axis):
    """Forward difference with a zero last row/column (OBGCCriterion.lua:67-68, 91-92;
    SmoothnessCriterion.lua:45-46): g[i] = a[i+1]-a[i], g[last] = 0."""
    g = np.zeros_like(a)
    if axis == 2:
        g[:, :, :-1, :] = a[:, :, 1:, :] - a[:, :, :-1, :]
    else:
        g[:, :, :, :-1] = a[:, :, :, 1:] - a[:, :, :, :-1]
    return g


class OBCriterionOracle:
    """OBCC (``gradient_terms=False``) and OBGCC (``True``).

    inputs: flow (B,2,h,w), bflow or None, occ (B,2,h,w), warped: list of F-1 (B,C,h,w),
    target (B,C,h,w).  Fields mirror the Lua criterion fields."""

    def __init__(self, gradient_terms=False, penalty=None, F=3, past_flow=False,
                 pwc_flow_scaling=1.0, penalty_out=1.0, size_average=True,
                 alpha=1.0, beta=1.0, gamma=1.0, grad_check=False, dtype=np.float64):
        self.gradient_terms = gradient_terms
        self.p = penalty or QuadraticPenalty()
        self.F = F
        self.past_flow = past_flow
        self.pwc_flow_scaling = pwc_flow_scaling
        self.penalty_out = penalty_out
        self.sizeAverage = size_average
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.gradCheck = grad_check
        self.dtype = dtype

    def _masks(self, flow, bflow):
        flows = (flow, bflow)
        masks = []
        for (_, fi, k) in _frame_roles(self.F, self.past_flow):
            if self.gradCheck:
                masks.append(None)
            else:
                masks.append(out_of_image_mask(flows[fi], k, self.pwc_flow_scaling))
        return masks

    def forward(self, flow, bflow, occ, warped, target):
        """OBCCriterion.lua:36-119 / OBGCCriterion.lua:39-149.  Note alpha is NOT applied in
        the forward (OBGCCriterion.lua:97)."""
        dt = self.dtype
        occ = np.asarray(occ).astype(dt)
        tgt = np.asarray(target).astype(dt)
        B, C, h, w = tgt.shape
        masks = self._masks(flow, bflow)
        acc = np.zeros((B, h, w), dtype=dt)
        if self.gradient_terms:
            tgy, tgx = _fwd_diff(tgt, 2), _fwd_diff(tgt, 3)
        for (oc, _, _), img, m in zip(_frame_roles(self.F, self.past_flow), warped, masks):
            img = np.asarray(img).astype(dt)
            tmp = self.p.apply(img - tgt).sum(axis=1)
            if self.gradient_terms:
                tmp = tmp + self.p.apply(_fwd_diff(img, 3) - tgx).sum(axis=1) * dt(self.beta)
                tmp = tmp + self.p.apply(_fwd_diff(img, 2) - tgy).sum(axis=1) * dt(self.gamma)
            tmp = tmp * occ[:, oc]
            if m is not None:
                mf = m.astype(dt)
                tmp = tmp * mf + (dt(1) - mf) * dt(self.penalty_out)
            acc += tmp
        out = float(acc.sum(dtype=np.float64)) / (C * (self.F - 1))
        if self.sizeAverage:
            out *= 1.0 / (B * h * w)
        return out

    def backward(self, flow, bflow, occ, warped, target):
        """OBCCriterion.lua:121-240 / OBGCCriterion.lua:151-300.

        Returns (gradOcc (B,2,h,w), [gradWarp_f ...]).  Quirks kept: the out-of-image
        penalty is added to the occlusion gradient (Q7); for OBGCC the occlusion gradient
        uses the +/- shifted combination of p(.) values (Q6) and alpha only appears here."""
        dt = self.dtype
        occ = np.asarray(occ).astype(dt)
        tgt = np.asarray(target).astype(dt)
        B, C, h, w = tgt.shape
        masks = self._masks(flow, bflow)
        g_occ = np.zeros((B, 2, h, w), dtype=dt)
        g_warp = []
        norm = dt(1.0 / (C * (self.F - 1)))
        if self.sizeAverage:
            norm = norm * dt(1.0 / (B * h * w))
        if self.gradient_terms:
            tgy, tgx = _fwd_diff(tgt, 2), _fwd_diff(tgt, 3)
        for (oc, _, _), img, m in zip(_frame_roles(self.F, self.past_flow), warped, masks):
            img = np.asarray(img).astype(dt)
            d = img - tgt
            if not self.gradient_terms:
                gi = self.p.der(d)
                buf = self.p.apply(d).sum(axis=1)
            else:
                dgy = _fwd_diff(img, 2) - tgy
                dgx = _fwd_diff(img, 3) - tgx
                gi = self.p.der(d) * dt(self.alpha)
                t = self.p.der(dgy) * dt(self.gamma)
                gi = gi - t
                gi[:, :, 1:, :] += t[:, :, :-1, :]
                t = self.p.der(dgx) * dt(self.beta)
                gi = gi - t
                gi[:, :, :, 1:] += t[:, :, :, :-1]
                buf = self.p.apply(d).sum(axis=1) * dt(self.alpha)
                t = self.p.apply(dgy).sum(axis=1) * dt(self.gamma)
                buf = buf - t
                buf[:, 1:, :] += t[:, :-1, :]
                t = self.p.apply(dgx).sum(axis=1) * dt(self.beta)
                buf = buf - t
                buf[:, :, 1:] += t[:, :, :-1]
            if m is not None:
                mf = m.astype(dt)
                buf = buf * mf + (dt(1) - mf) * dt(self.penalty_out)
                gi = gi * mf[:, None]
            g_occ[:, oc] += buf
            gi = gi * occ[:, oc][:, None]
            g_warp.append(gi * norm)
        g_occ *= norm
        return g_occ, g_warp


# ----------------------------------------------------------------------------------------
# SmoothnessCriterion / SecondOrderSmoothnessCriterion
# ----------------------------------------------------------------------------------------

class _TorchStorageTensor:
    """Just enough of TH's tensor/storage model to replay SmoothnessCriterion.lua:49-56
    literally: a tensor is (storage, offset, size, stride); ``cadd`` first does
    ``resizeAs(result, src1)``, which for a size mismatch re-lays the result out contiguously
    from its storage offset and grows the shared storage (TH ``THTensor_(resizeNd)``)."""

    def __init__(self, storage, offset, size, stride):
        self.storage, self.offset, self.size, self.stride = storage, offset, tuple(size), tuple(stride)

    @staticmethod
    def zeros(size, dtype):
        n = int(np.prod(size))
        stride = _contig_strides(size)
        return _TorchStorageTensor([np.zeros(n, dtype=dtype)], 0, size, stride)

    def narrow(self, dim, start, length):
        size = list(self.size)
        size[dim] = length
        return _TorchStorageTensor(self.storage, self.offset + start * self.stride[dim], size, self.stride)

    def view_np(self):
        itemsize = self.storage[0].itemsize
        return np.lib.stride_tricks.as_strided(
            self.storage[0][self.offset:], shape=self.size,
            strides=tuple(s * itemsize for s in self.stride), writeable=True)

    def cadd_overwrite(self, a, b_scaled):
        """r:add(a, v, b) with the TH resizeAs semantics; ``b_scaled`` is already v*b."""
        if tuple(a.shape) != self.size:
            self.size = tuple(a.shape)
            self.stride = _contig_strides(self.size)
            need = self.offset + int(np.prod(self.size))
            if need > self.storage[0].size:
                grown = np.zeros(need, dtype=self.storage[0].dtype)
                grown[:self.storage[0].size] = self.storage[0]
                self.storage[0] = grown
        self.view_np()[...] = a + b_scaled


def _contig_strides(size):
    st, acc = [], 1
    for s in reversed(size):
        st.append(acc)
        acc *= s
    return tuple(reversed(st))


def smooth1_weight_inputs_literal(inp_shape, target, dtype=np.float64):
    """Replays SmoothnessCriterion.lua:49-56 on the mini storage model and returns the
    (B,Cin,h,w) arrays ``igy``/``igx`` exactly as the following ``torch.abs(igy)`` sees
    them."""
    tgt = np.asarray(target).astype(dtype)
    B, Cin, h, w = inp_shape
    igy = _TorchStorageTensor.zeros(inp_shape, dtype)
    igx = _TorchStorageTensor.zeros(inp_shape, dtype)
    igy.narrow(2, 0, h - 1).cadd_overwrite(tgt[:, :, 1:, :], -tgt[:, :, :-1, :])
    igx.narrow(3, 0, w - 1).cadd_overwrite(tgt[:, :, :, 1:], -tgt[:, :, :, :-1])
    return igy.view_np().copy(), igx.view_np().copy()


def smooth1_weight_inputs(inp_shape, target, alias=True, dtype=np.float64):
    """``igy``/``igx`` of SmoothnessCriterion.lua:49-56 in closed form.

    alias=True (parity default, Q9): when the input's channel count differs from the
    target's, the narrowed views are re-laid out contiguously, so ``igy`` reads the first
    B*Cin*h*w elements of the contiguous (B,Ct,h-1,w) difference array (zero beyond its
    end) and ``igx`` those of the contiguous (B,Ct,h,w-1) array.  When the channel counts
    agree no resize happens and the natural in-place result is obtained.
    alias=False: the evidently intended weights -- forward differences of the target with a
    zero last row/column, all Ct channels."""
    tgt = np.asarray(target).astype(dtype)
    B, Cin, h, w = inp_shape
    Ct = tgt.shape[1]
    dy = tgt[:, :, 1:, :] - tgt[:, :, :-1, :]
    dx = tgt[:, :, :, 1:] - tgt[:, :, :, :-1]
    if not alias or Ct == Cin:
        igy = np.zeros((B, Ct, h, w), dtype=dtype)
        igx = np.zeros((B, Ct, h, w), dtype=dtype)
        igy[:, :, :-1, :] = dy
        igx[:, :, :, :-1] = dx
        return igy, igx
    n = B * Cin * h * w

    def take(flat):
        buf = np.zeros(n, dtype=dtype)
        m = min(n, flat.size)
        buf[:m] = flat[:m]
        return buf.reshape(B, Cin, h, w)

    return take(np.ascontiguousarray(dy).ravel()), take(np.ascontiguousarray(dx).ravel())


class SmoothnessOracle:
    """SmoothnessCriterion (order=1) / SecondOrderSmoothnessCriterion (order=2)."""

    def __init__(self, order=1, penalty=None, cs=20.0, size_average=True, alias=True,
                 dtype=np.float64):
        self.order, self.p, self.cs = order, penalty or QuadraticPenalty(), cs
        self.sizeAverage, self.alias, self.dtype = size_average, alias, dtype

    def _g_and_w(self, inp, target):
        dt = self.dtype
        x = np.asarray(inp).astype(dt)
        tgt = np.asarray(target).astype(dt)
        B, Cin, h, w = x.shape
        if self.order == 1:
            # SmoothnessCriterion.lua:45-46, 49-59
            gy, gx = _fwd_diff(x, 2), _fwd_diff(x, 3)
            igy, igx = smooth1_weight_inputs(x.shape, target, self.alias, dt)
            wy = np.exp(dt(-self.cs) * np.abs(igy).mean(axis=1))
            wx = np.exp(dt(-self.cs) * np.abs(igx).mean(axis=1))
        else:
            # SecondOrderSmoothnessCriterion.lua:45-46: (2 I[i] - I[i-1]) - I[i+1], interior only
            gy = np.zeros_like(x)
            gx = np.zeros_like(x)
            gy[:, :, 1:-1, :] = (dt(2) * x[:, :, 1:-1, :] - x[:, :, :-2, :]) - x[:, :, 2:, :]
            gx[:, :, :, 1:-1] = (dt(2) * x[:, :, :, 1:-1] - x[:, :, :, :-2]) - x[:, :, :, 2:]
            # :49-58: igy[1:] += mean|T[i]-T[i-1]| ; igy[1:-1] += mean|T[i]-T[i+1]|
            igy = np.zeros((B, h, w), dtype=dt)
            igx = np.zeros((B, h, w), dtype=dt)
            igy[:, 1:, :] += np.abs(tgt[:, :, 1:, :] - tgt[:, :, :-1, :]).mean(axis=1)
            igx[:, :, 1:] += np.abs(tgt[:, :, :, 1:] - tgt[:, :, :, :-1]).mean(axis=1)
            igy[:, 1:-1, :] += np.abs(tgt[:, :, 1:-1, :] - tgt[:, :, 2:, :]).mean(axis=1)
            igx[:, :, 1:-1] += np.abs(tgt[:, :, :, 1:-1] - tgt[:, :, :, 2:]).mean(axis=1)
            wy = np.exp(dt(-self.cs) * igy)
            wx = np.exp(dt(-self.cs) * igx)
        return gy, gx, wy[:, None], wx[:, None]

    def forward(self, inp, target):
        gy, gx, wy, wx = self._g_and_w(inp, target)
        buf = self.p.apply(gx) * wx + self.p.apply(gy) * wy
        out = float(buf.sum(dtype=np.float64))
        if self.sizeAverage:
            out *= 1.0 / buf.size
        return out

    def backward(self, inp, target):
        dt = self.dtype
        gy, gx, wy, wx = self._g_and_w(inp, target)
        Gy = self.p.der(gy) * wy
        Gx = self.p.der(gx) * wx
        g = np.zeros_like(gy)
        if self.order == 1:
            # SmoothnessCriterion.lua:85-103: -Gx + shift(Gx) - Gy + shift(Gy)
            g = -Gx
            g[:, :, :, 1:] += Gx[:, :, :, :-1]
            g = g - Gy
            g[:, :, 1:, :] += Gy[:, :, :-1, :]
        else:
            # SecondOrderSmoothnessCriterion.lua:90-97
            g[:, :, 1:-1, :] += dt(2) * Gy[:, :, 1:-1, :]
            g[:, :, :, 1:-1] += dt(2) * Gx[:, :, :, 1:-1]
            g[:, :, :-2, :] -= Gy[:, :, 1:-1, :]
            g[:, :, :, :-2] -= Gx[:, :, :, 1:-1]
            g[:, :, 2:, :] -= Gy[:, :, 1:-1, :]
            g[:, :, :, 2:] -= Gx[:, :, :, 1:-1]
        if self.sizeAverage:
            g = g * dt(1.0 / g.size)
        return g


# ----------------------------------------------------------------------------------------
# ConstVelCriterion / OcclusionPriorCriterion
# ----------------------------------------------------------------------------------------

def constvel_forward(f, b, size_average=True, dtype=np.float64):
    """ConstVelCriterion.lua:29-46: sum_px ||f-b||_2, times 1/nElement(f) (= 1/(B*2*h*w))."""
    f = np.asarray(f).astype(dtype)
    b = np.asarray(b).astype(dtype)
    out = float(np.sqrt(((f - b) ** 2).sum(axis=1)).sum(dtype=np.float64))
    if size_average:
        out *= 1.0 / f.size
    return out


def constvel_backward(f, b, size_average=True, dtype=np.float64):
    """ConstVelCriterion.lua:48-74: +-(f-b)/(||f-b||+1e-12), divided by npixels = B*h*w (the
    forward and backward normalisers differ by the channel count)."""
    f = np.asarray(f).astype(dtype)
    b = np.asarray(b).astype(dtype)
    den = np.sqrt(((f - b) ** 2).sum(axis=1, keepdims=True)) + dtype(1e-12)
    g1 = (f - b) / den
    g2 = (b - f) / den
    if size_average:
        npix = dtype(f.size // f.shape[1])
        g1, g2 = g1 / npix, g2 / npix
    return g1, g2


def occprior_forward(occ, size_average=True, penalty=1.0, dtype=np.float64):
    """OcclusionPriorCriterion.lua:28-49 (2- and 3-channel branches)."""
    o = np.asarray(occ).astype(dtype)
    B, C, h, w = o.shape
    if C == 3:
        v = (dtype(1) - o[:, 1]) * (o[:, 0] + o[:, 2]) * dtype(penalty) * dtype(0.05)
    else:
        v = (dtype(1) - o[:, 0] * o[:, 1]) * dtype(penalty)
    out = float(v.sum(dtype=np.float64))
    if size_average:
        out *= 1.0 / (B * h * w)
    return out


def occprior_backward(occ, size_average=True, penalty=1.0, dtype=np.float64):
    """OcclusionPriorCriterion.lua:51-73 -- not the analytic derivative."""
    o = np.asarray(occ).astype(dtype)
    B, C, h, w = o.shape
    g = o.copy()
    if C == 3:
        g[:, 0] = (dtype(1) - o[:, 1]) * dtype(penalty) * dtype(0.05)
        g[:, 1] = -(o[:, 0] + o[:, 2]) * dtype(penalty) * dtype(0.05)
        g[:, 2] = (dtype(1) - o[:, 1]) * dtype(penalty) * dtype(0.05)
    else:
        g[:, 0] = (dtype(1) - o[:, 1]) * dtype(penalty)
        g[:, 1] = (dtype(1) - o[:, 0]) * dtype(penalty)
    if size_average:
        g = g * dtype(1.0 / (B * h * w))
    return g


# ----------------------------------------------------------------------------------------
# tolerance helper shared by the parity tests
# ----------------------------------------------------------------------------------------

def rel_err(a, b):
    """max |a-b| / max(|b|, rms(b)) -- the 1e-4 relative criterion of north_star with the
    scale floor of SURVEY 8c (cost volumes and gradients cross zero)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if b.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b), np.sqrt(np.mean(b * b)) + 1e-30)
    return float(np.max(np.abs(a - b) / scale))
