"""CPU tests that pin the oracle (parity is unpinned by the reference's own tests).

  * two independent restatements agree (literal Lua-loop transcription vs closed form; vectorised
    vs scalar sampler; literal Torch-storage replay vs closed-form aliasing; numpy vs C library);
  * finite-difference gradient checks in the spirit of the reference's commented-out Jacobian
    tests (models/CostVolMulti.lua:192-223: ws 5, bs 1, c 2, f 7, precision 1e-5;
    extras/stnbhwd/test.lua:47-120);
  * derived known-answer cases (delta images of models/CostVolMulti.lua:225-254; identity /
    integer-shift flows for the sampler);
  * the documented quirks Q1-Q15 behave as documented.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import b2f_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RNG = np.random.default_rng(2)  # the reference's -manualSeed default (opts.lua:27)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------------------
# cost volume
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("win,F,fwd", [(9, 2, True), (9, 2, False), (5, 3, True), (5, 3, False), (3, 4, True)])
def test_costvol_literal_vs_closed(win, F, fwd):
    n = (win - 1) // 2
    h, w = n * (F - 1) + 3, n * (F - 1) + 5   # Torch7 would raise on inverted ranges below this
    frames = [f32(RNG.standard_normal((2, 3, h, w))) for _ in range(F)]
    a = o.costvol_forward_lua(frames, win, fwd)
    b = o.costvol_forward(frames, win, fwd)
    assert np.abs(a - b).max() < 1e-12


def test_costvol_channel_order_and_direction():
    """Q4: channel index is x-major and the source pixel is p - q (fwd) / p + q (bwd)."""
    h = w = 12
    ref = np.zeros((1, 1, h, w), np.float32)
    frm = np.zeros((1, 1, h, w), np.float32)
    ref[0, 0, 6, 5] = 1.0          # (y, x) = (6, 5)
    frm[0, 0, 4, 8] = 1.0          # source pixel: y - qy = 4 -> qy = 2 ; x - qx = 8 -> qx = -3
    out = o.costvol_forward([ref, frm], 9, True)
    nz = np.argwhere(out != 0)
    assert nz.tolist() == [[0, (-3 + 4) * 9 + (2 + 4), 6, 5]]
    assert out[tuple(nz[0])] == 1.0   # C*(F-1) = 1
    out = o.costvol_forward([ref, frm], 9, False)   # mirrored window: source = p + q
    nz = np.argwhere(out != 0)
    assert nz.tolist() == [[0, (3 + 4) * 9 + (-2 + 4), 6, 5]]


def test_costvol_delta_images_kat():
    """The commented toy test of models/CostVolMulti.lua:225-254: 6x6 delta images moving by
    +-1 px per frame, win 5, 3 frames.  Frame f (displacement multiplier m) sits at (3+m, 3+m)
    (1-based) in the future table and (3-m, 3-m) in the past table; with fwd=true and source
    p - m q the future deltas are found at q = (-1,-1); with fwd=false and the past table the
    same channel fires.  The value is 2 hits / (C (F-1)) = 1."""
    def delta(i):
        img = np.zeros((1, 1, 6, 6), np.float32)
        img[0, 0, i - 1, i - 1] = 1
        return img
    future = [delta(3), delta(4), delta(5)]
    past = [delta(3), delta(2), delta(1)]
    for frames, fwd in ((future, True), (past, False)):
        out = o.costvol_forward_lua(frames, 5, fwd)
        ch = (-1 + 2) * 5 + (-1 + 2)
        assert out[0, ch, 2, 2] == 1.0
        assert np.count_nonzero(out) == 1


def test_costvol_constant_normaliser_at_borders():
    """The normaliser is C*(F-1) everywhere, also where window terms were dropped (:100)."""
    ref = np.ones((1, 4, 10, 10), np.float32)
    out = o.costvol_forward([ref, ref], 9, True)
    assert out[0, 40, 5, 5] == 1.0            # centre displacement
    assert out[0, 0, 0, 0] == 1.0             # q = (-4,-4): source (4,4), in range
    assert out[0, 80, 0, 0] == 0.0            # q = (+4,+4): source (-4,-4), dropped
    assert out[0, 80, 9, 9] == 1.0


@pytest.mark.parametrize("fwd", [True, False])
def test_costvol_backward_is_adjoint(fwd):
    """<J dx, go> == <dx, J^T go> for the bilinear map (exact up to fp64 rounding); the
    reference's commented Jacobian test used ws 5, bs 1, c 2, f 7."""
    win, F = 5, 3
    frames = [RNG.standard_normal((1, 2, 9, 11)) for _ in range(F)]
    go = RNG.standard_normal((1, win * win, 9, 11))
    grads = o.costvol_backward(frames, go, win, fwd)
    for k in range(F):
        d = RNG.standard_normal(frames[k].shape)
        eps = 1e-6
        fp = [f.copy() for f in frames]
        fm = [f.copy() for f in frames]
        fp[k] += eps * d
        fm[k] -= eps * d
        num = ((o.costvol_forward(fp, win, fwd) - o.costvol_forward(fm, win, fwd)) * go).sum() / (2 * eps)
        ana = (grads[k] * d).sum()
        assert abs(num - ana) < 1e-6 * max(1.0, abs(ana))


# ---------------------------------------------------------------------------------------
# sampler
# ---------------------------------------------------------------------------------------

def test_warp_vectorised_vs_scalar():
    img = f32(RNG.standard_normal((2, 6, 7, 3)))
    grid = f32(RNG.standard_normal((2, 6, 7, 2)) * 3)
    assert np.abs(o.warp_forward(img, grid) - o.warp_forward_loops(img, grid)).max() < 1e-12


def test_warp_identity_and_integer_shift():
    img = f32(RNG.standard_normal((1, 5, 8, 4)))
    zero = np.zeros((1, 5, 8, 2), np.float32)
    assert np.array_equal(o.warp_forward(img, zero, np.float32), img)
    g = zero.copy()
    g[..., 0] = 2.0   # channel 0 is x (Q2): out[y, x] = img[y, min(x+2, W-1)], clamped (Q3)
    out = o.warp_forward(img, g, np.float32)
    assert np.array_equal(out[:, :, :6], img[:, :, 2:])
    assert np.array_equal(out[:, :, 6:], np.repeat(img[:, :, 7:8], 2, axis=2))
    g = zero.copy()
    g[..., 1] = -1.0  # channel 1 is y
    out = o.warp_forward(img, g, np.float32)
    assert np.array_equal(out[:, 1:], img[:, :-1])
    assert np.array_equal(out[:, 0], img[:, 0])


def test_warp_tap_at_W_reads_zero_but_weight_is_zero():
    """Q3: at xc == W-1 exactly the right tap is out of range; its weight is 0 anyway."""
    img = f32(RNG.standard_normal((1, 3, 4, 2)))
    g = np.zeros((1, 3, 4, 2), np.float32)
    g[..., 0] = 100.0
    out = o.warp_forward(img, g, np.float32)
    assert np.array_equal(out, np.repeat(img[:, :, 3:4], 4, axis=2))


def test_warp_backward_grad_img_is_adjoint_and_grad_grid_is_derivative():
    img = RNG.standard_normal((1, 6, 7, 3)).astype(np.float32)
    grid = (RNG.uniform(-1.5, 1.5, (1, 6, 7, 2))).astype(np.float32)
    go = RNG.standard_normal((1, 6, 7, 3)).astype(np.float32)
    g_img, g_grid = o.warp_backward(img, grid, go)
    # image gradient: exact adjoint of the (linear in img) forward
    d = RNG.standard_normal(img.shape).astype(np.float32)
    lhs = (o.warp_forward(d, grid) * go).sum()
    rhs = (g_img * d).sum()
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))
    # flow gradient: derivative where no clamp / cell boundary is crossed
    xs = np.arange(7)[None, None, :] + grid[..., 0]
    ys = np.arange(6)[None, :, None] + grid[..., 1]
    safe = ((xs > 0.05) & (xs < 5.95) & (ys > 0.05) & (ys < 4.95)
            & (np.abs(xs - np.round(xs)) > 0.02) & (np.abs(ys - np.round(ys)) > 0.02))
    eps = np.float32(2.0 ** -6)   # bilinear is piecewise linear: exact inside a cell, fp32 geometry noise / eps
    for ch in range(2):
        gp, gm = grid.copy(), grid.copy()
        gp[..., ch] += eps
        gm[..., ch] -= eps
        num = ((o.warp_forward(img, gp) - o.warp_forward(img, gm)) * go).sum(axis=-1) / (gp[..., ch] - gm[..., ch])
        assert np.abs(num - g_grid[..., ch])[safe].max() < 1e-4
    assert safe.sum() > 10


def test_warp_no_clamp_derivative():
    """Q3: clamped pixels still receive a flow gradient (the reference does not zero it)."""
    img = f32(np.arange(2 * 4 * 1).reshape(1, 2, 4, 1))
    g = np.zeros((1, 2, 4, 2), np.float32)
    g[..., 0] = -50.0     # every sample clamps to x = 0
    go = np.ones((1, 2, 4, 1), np.float32)
    _, gg = o.warp_backward(img, g, go)
    # xi = 0, wx = 1, wy = 1: gradGrid.x = -wy*D_TL + wy*D_TR = img[y,1]-img[y,0] = 1
    assert np.allclose(gg[..., 0], 1.0)


def test_warp_only_grid_matches_full():
    img = f32(RNG.standard_normal((2, 5, 6, 4)))
    grid = f32(RNG.standard_normal((2, 5, 6, 2)) * 2)
    go = f32(RNG.standard_normal((2, 5, 6, 4)))
    gi, gg = o.warp_backward(img, grid, go)
    gi2, gg2 = o.warp_backward(img, grid, go, only_grid=True)
    assert gi2 is None and np.array_equal(gg, gg2)


# ---------------------------------------------------------------------------------------
# C restatement vs numpy
# ---------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def cpu_lib():
    path = os.path.join(ROOT, "oracle", "c", "libb2f_cpu.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build_oracle()
    lib = C.CDLL(path)
    return lib


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("fwd", [1, 0])
def test_c_costvol_matches_numpy(cpu_lib, fwd):
    B, Cn, h, w, win, F = 2, 5, 11, 13, 9, 2
    frames = [f32(RNG.standard_normal((B, Cn, h, w))) for _ in range(F)]
    ptrs = (C.c_void_p * F)(*[f.ctypes.data for f in frames])
    out = np.empty((B, win * win, h, w), np.float32)
    assert cpu_lib.b2fcpu_costvol_forward(ptrs, F, B, Cn, h, w, win, fwd, _fp(out)) == 0
    ref = o.costvol_forward(frames, win, bool(fwd))
    assert o.rel_err(out, ref) < 1e-5
    # backward through a batch-strided gradOut (narrow of a 2*81-channel buffer)
    wide = f32(RNG.standard_normal((B, 2 * win * win, h, w)))
    go = wide[:, win * win:]
    grads = [np.empty_like(f) for f in frames]
    gptrs = (C.c_void_p * F)(*[g.ctypes.data for g in grads])
    cpu_lib.b2fcpu_costvol_backward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    assert cpu_lib.b2fcpu_costvol_backward(ptrs, F, B, Cn, h, w, win, fwd, C.c_void_p(go.ctypes.data),
                                           wide.strides[0] // 4, gptrs) == 0
    refg = o.costvol_backward(frames, go, win, bool(fwd))
    for a, b in zip(grads, refg):
        assert o.rel_err(a, b) < 1e-5


def test_c_warp_matches_numpy(cpu_lib):
    B, H, W, Cn = 2, 9, 10, 3
    img = f32(RNG.standard_normal((B, H, W, Cn)))
    grid = f32(RNG.standard_normal((B, H, W, 2)) * 4)
    go = f32(RNG.standard_normal((B, H, W, Cn)))
    out = np.empty_like(img)
    assert cpu_lib.b2fcpu_warp_forward(_fp(img), _fp(grid), _fp(out), B, H, W, Cn, H, W) == 0
    assert o.rel_err(out, o.warp_forward(img, grid)) < 1e-5
    gi = np.zeros_like(img)
    gg = np.empty_like(grid)
    assert cpu_lib.b2fcpu_warp_backward(_fp(img), _fp(grid), _fp(go), _fp(gi), _fp(gg), B, H, W, Cn, H, W) == 0
    rgi, rgg = o.warp_backward(img, grid, go)
    assert o.rel_err(gi, rgi) < 1e-5 and o.rel_err(gg, rgg) < 1e-5


# ---------------------------------------------------------------------------------------
# criterions
# ---------------------------------------------------------------------------------------

def _ob_inputs(B=2, Cn=3, h=7, w=9, flow_sigma=0.3):
    flow = f32(RNG.standard_normal((B, 2, h, w)) * flow_sigma)
    bflow = f32(RNG.standard_normal((B, 2, h, w)) * flow_sigma)
    e = np.exp(RNG.standard_normal((B, 2, h, w)))
    occ = f32(e / e.sum(axis=1, keepdims=True))
    warped = [f32(RNG.uniform(-2.1, 2.6, (B, Cn, h, w))) for _ in range(2)]
    target = f32(RNG.uniform(-2.1, 2.6, (B, Cn, h, w)))
    return flow, bflow, occ, warped, target


@pytest.mark.parametrize("gt", [False, True])
@pytest.mark.parametrize("pen", [o.PEN_QUADRATIC, o.PEN_L1, o.PEN_LORENTZIAN])
def test_ob_warp_gradient_is_derivative_of_loss(gt, pen):
    """With alpha = 1 the warped-frame gradient is the true derivative of the forward energy
    (the reference's gradCheck mode: no mask).  For OBGCC this only holds for a spatially constant
    occlusion map: the reference multiplies the shifted gradient-term derivatives by occ (and the
    mask) of the RECEIVING pixel (OBGCCriterion.lua:200-212, 257, 290), not of the pixel whose
    energy they belong to -- kept as is."""
    flow, bflow, occ, warped, target = _ob_inputs()
    if gt:
        occ = np.full_like(occ, 0.5)
    crit = o.OBCriterionOracle(gt, o.make_penalty(pen), grad_check=True, size_average=False,
                               beta=0.7, gamma=1.3)
    _, g_warp = crit.backward(flow, None, occ, warped, target)
    eps = 1e-4
    for k in range(2):
        d = RNG.standard_normal(warped[k].shape)
        wp = [x.astype(np.float64) for x in warped]
        wm = [x.astype(np.float64) for x in warped]
        wp[k] = wp[k] + eps * d
        wm[k] = wm[k] - eps * d
        # the oracle casts inputs to fp32 first; perturb in a way fp32 can represent
        num = (_ob_forward64(crit, flow, occ, wp, target) - _ob_forward64(crit, flow, occ, wm, target)) / (2 * eps)
        ana = (g_warp[k] * d).sum()
        assert abs(num - ana) < 2e-3 * max(1.0, abs(ana))


def _ob_forward64(crit, flow, occ, warped64, target):
    """Forward of the OB oracle on float64 warped frames (bypasses the fp32 input cast)."""
    dt = np.float64
    tgt = target.astype(dt)
    occ = occ.astype(dt)
    acc = 0.0
    for (oc, _, _), img in zip(o._frame_roles(3, False), warped64):
        tmp = crit.p.apply(img - tgt).sum(axis=1)
        if crit.gradient_terms:
            tmp = tmp + crit.p.apply(o._fwd_diff(img, 3) - o._fwd_diff(tgt, 3)).sum(axis=1) * crit.beta
            tmp = tmp + crit.p.apply(o._fwd_diff(img, 2) - o._fwd_diff(tgt, 2)).sum(axis=1) * crit.gamma
        acc += (tmp * occ[:, oc]).sum()
    return acc / (tgt.shape[1] * 2)


def test_ob_mask_penalty_and_occ_gradient_quirks():
    """Q7: masked-out pixels contribute penalty_out to the loss AND to the occlusion gradient;
    their warped-frame gradient is zero.  Q14: mask thresholds are inclusive at 1 and w/h."""
    B, Cn, h, w = 1, 3, 4, 6
    flow = np.zeros((B, 2, h, w), np.float32)
    flow[0, 0, 1, 2] = 10.0    # future frame (k=+1): x = 3 + 10*1 > w -> out; past (k=-1): 3-10 < 1 -> out
    flow[0, 0, 2, 5] = 0.0     # x = 6 == w -> in (inclusive)
    occ = np.full((B, 2, h, w), 0.5, np.float32)
    warped = [f32(RNG.standard_normal((B, Cn, h, w))) for _ in range(2)]
    target = f32(RNG.standard_normal((B, Cn, h, w)))
    crit = o.OBCriterionOracle(False, o.L1Penalty(), size_average=False, penalty_out=0.37)
    m = o.out_of_image_mask(flow, 1, 1.0)
    assert not m[0, 1, 2] and m.sum() == h * w - 1
    g_occ, g_warp = crit.backward(flow, None, occ, warped, target)
    norm = 1.0 / (Cn * 2)
    assert np.isclose(g_occ[0, 0, 1, 2], 0.37 * norm) and np.isclose(g_occ[0, 1, 1, 2], 0.37 * norm)
    assert np.all(g_warp[0][0, :, 1, 2] == 0) and np.all(g_warp[1][0, :, 1, 2] == 0)
    # loss = sum over frames of (energy*occ masked + penalty)
    loss = crit.forward(flow, None, occ, warped, target)
    crit_nomask = o.OBCriterionOracle(False, o.L1Penalty(), size_average=False, grad_check=True)
    full = crit_nomask.forward(flow, None, occ, warped, target)
    e = [o.L1Penalty().apply((wv - target).astype(np.float64)).sum(axis=1)[0, 1, 2] * 0.5 for wv in warped]
    assert np.isclose(loss, full - sum(e) * norm + 2 * 0.37 * norm)


def test_ob_mask_rounding_order():
    """Q14: three separately rounded fp32 operations.  1-based coordinate 1 plus
    fl(fl(-1*flow)*scale) = -fl(20*flow); with flow = 1.4901161e-09 (2^-29.32..) the product is
    far below ulp(1)/2 so tcoord == 1 -> inside; a double-precision evaluation would give < 1."""
    flow = np.zeros((1, 2, 2, 2), np.float32)
    flow[0, 0, 0, 0] = np.float32(1.4901161e-09)
    m = o.out_of_image_mask(flow, -1, 20.0)
    assert m[0, 0, 0]
    assert 1.0 + (-1.0 * float(flow[0, 0, 0, 0])) * 20.0 < 1.0


def test_ob_past_flow_uses_second_flow_for_past_frame():
    flow, bflow, occ, warped, target = _ob_inputs(flow_sigma=3.0)
    a = o.OBCriterionOracle(False, o.L1Penalty(), past_flow=True, pwc_flow_scaling=2.5, size_average=False)
    la = a.forward(flow, bflow, occ, warped, target)
    lb = a.forward(flow, flow, occ, warped, target)
    assert la != lb


def test_obgcc_alpha_backward_only_and_occ_grad_not_derivative():
    """Q5 / Q6."""
    flow, bflow, occ, warped, target = _ob_inputs()
    c1 = o.OBCriterionOracle(True, o.L1Penalty(), alpha=0.0, grad_check=True)
    c2 = o.OBCriterionOracle(True, o.L1Penalty(), alpha=1.0, grad_check=True)
    assert c1.forward(flow, None, occ, warped, target) == c2.forward(flow, None, occ, warped, target)
    g1, _ = c1.backward(flow, None, occ, warped, target)
    g2, _ = c2.backward(flow, None, occ, warped, target)
    assert np.abs(g1 - g2).max() > 0


@pytest.mark.parametrize("shape", [(2, 2, 5, 6, 3), (3, 2, 4, 4, 3), (1, 2, 2, 9, 3), (2, 2, 3, 3, 3), (2, 3, 5, 4, 3)])
def test_smooth1_alias_closed_form_matches_literal_replay(shape):
    """Q9: the closed-form aliased weights equal a literal replay of the Torch storage model."""
    B, Cin, h, w, Ct = shape
    target = f32(RNG.standard_normal((B, Ct, h, w)))
    ly, lx = o.smooth1_weight_inputs_literal((B, Cin, h, w), target)
    cy, cx = o.smooth1_weight_inputs((B, Cin, h, w), target, alias=True)
    assert np.array_equal(ly, cy) and np.array_equal(lx, cx)


def test_smooth1_alias_batch0_row_structure():
    """The survey's description of Q9: for batch 0, igy[0, j=0] is |dy R| and igy[0, j=1, y] is
    dy of R/G shifted (rows of the (h-1)-row difference array re-read with h-row pitch)."""
    B, h, w = 2, 6, 5
    target = f32(RNG.standard_normal((B, 3, h, w)))
    igy, _ = o.smooth1_weight_inputs((B, 2, h, w), target, alias=True)
    t64 = target.astype(np.float64)
    dy = t64[:, :, 1:] - t64[:, :, :-1]
    assert np.array_equal(igy[0, 0, :h - 1], dy[0, 0])          # rows 0..h-2 of channel R
    assert np.array_equal(igy[0, 0, h - 1], dy[0, 1, 0])        # last row spills into G's first row
    assert np.array_equal(igy[0, 1, :h - 2], dy[0, 1, 1:])      # second plane starts at G row 1


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("pen", [o.PEN_QUADRATIC, o.PEN_L1, o.PEN_LORENTZIAN])
def test_smoothness_gradient_is_derivative(order, pen):
    B, h, w = 2, 7, 8
    x = RNG.standard_normal((B, 2, h, w))
    target = f32(RNG.uniform(-2, 2, (B, 3, h, w)) * 0.05)
    crit = o.SmoothnessOracle(order, o.make_penalty(pen), size_average=False)
    g = crit.backward(x, target)
    d = RNG.standard_normal(x.shape)
    eps = 1e-6
    num = (crit.forward(x + eps * d, target) - crit.forward(x - eps * d, target)) / (2 * eps)
    ana = (g * d).sum()
    assert abs(num - ana) < 1e-5 * max(1.0, abs(ana))


def test_smoothness_counts_zero_difference_border():
    """Q8 / row a10: p_L1(0) = 1e-3 contributes on the zero last row / column."""
    x = np.zeros((1, 2, 3, 4), np.float32)
    t = np.zeros((1, 3, 3, 4), np.float32)
    loss = o.SmoothnessOracle(1, o.L1Penalty(), size_average=False).forward(x, t)
    assert np.isclose(loss, 2 * 3 * 4 * 2 * 1e-3)


def test_constvel_normalisers_differ():
    """Q11: forward is divided by nElement = B*2*h*w, backward by npixels = B*h*w."""
    f = f32(RNG.standard_normal((2, 2, 3, 4)))
    b = f32(RNG.standard_normal((2, 2, 3, 4)))
    l_avg = o.constvel_forward(f, b, True)
    l_sum = o.constvel_forward(f, b, False)
    assert np.isclose(l_avg * f.size, l_sum)
    g_avg, _ = o.constvel_backward(f, b, True)
    g_sum, g2 = o.constvel_backward(f, b, False)
    assert np.allclose(g_avg * (f.size // 2), g_sum) and np.allclose(g_sum, -g2)


def test_occprior_gradient_is_not_analytic():
    """Q15: 'grad' = (1-o2, 1-o1), not -(o2, o1)."""
    occ = f32(RNG.uniform(0, 1, (1, 2, 2, 2)))
    g = o.occprior_backward(occ, False)
    assert np.allclose(g[:, 0], 1 - occ[:, 1]) and np.allclose(g[:, 1], 1 - occ[:, 0])
    assert np.isclose(o.occprior_forward(occ, False), (1 - occ[:, 0] * occ[:, 1]).sum())
    occ3 = f32(RNG.uniform(0, 1, (1, 3, 2, 2)))
    assert np.isclose(o.occprior_forward(occ3, False), ((1 - occ3[:, 1]) * (occ3[:, 0] + occ3[:, 2])).sum() * 0.05)


def test_l1_penalty_ignores_alpha():
    """Q8: L1Penalty(0.38) behaves exactly like L1Penalty()."""
    x = np.linspace(-1, 1, 7)
    assert np.array_equal(o.L1Penalty(0.38).apply(x), o.L1Penalty().apply(x))
    assert np.isclose(o.L1Penalty().apply(np.zeros(1))[0], 1e-3)


# ---------------------------------------------------------------------------------------
# warpingUnit composition (models/pwc.lua:68-73 + MulConstant)
# ---------------------------------------------------------------------------------------

def test_warping_unit_composition_identity_adjoint_and_flow_gradient():
    r = np.random.default_rng(31)
    img = r.standard_normal((2, 5, 6, 7)).astype(np.float32)
    flow = (r.standard_normal((2, 2, 6, 7)) * 0.3).astype(np.float32)
    go = r.standard_normal(img.shape)
    # zero flow is the identity whatever the scale; channel 0 of the flow moves along x
    assert np.array_equal(o.warping_unit_forward(img, flow * 0, 3.0), img.astype(np.float64))
    shift = np.zeros_like(flow)
    shift[:, 0] = 0.5
    out = o.warping_unit_forward(img, shift, 4.0)                       # +2 px in x
    assert np.array_equal(out[..., :5], img[..., 2:].astype(np.float64))
    # gradImg is the adjoint of the (linear in img) forward
    gi, gf = o.warping_unit_backward(img, flow, 4.0, go)
    probe = r.standard_normal(img.shape).astype(np.float32)
    assert abs((o.warping_unit_forward(probe, flow, 4.0) * go).sum() - (probe * gi).sum()) < 1e-9
    # gradFlow carries MulConstant's factor: finite differences in network units
    for idx in ((0, 0, 2, 3), (1, 1, 4, 1)):
        eps = 1e-3
        f2, f1 = flow.copy(), flow.copy()
        f2[idx] += eps
        f1[idx] -= eps
        fd = ((o.warping_unit_forward(img, f2, 4.0) - o.warping_unit_forward(img, f1, 4.0)) * go).sum() / (2 * eps)
        assert abs(fd - gf[idx]) < 1e-3 * max(1.0, abs(fd))
    gi2, gf2 = o.warping_unit_backward(img, flow, 4.0, go, only_grid=True)
    assert gi2 is None and np.array_equal(gf2, gf)


# ---------------------------------------------------------------------------------------
# oracle/c/b2f_check64.c (the fast float64 checker of the full-size parity gates) pinned to the numpy oracle
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("B,Cn,h,w,win,F", [(2, 5, 7, 9, 9, 2), (1, 3, 6, 5, 5, 3), (2, 4, 4, 12, 3, 4), (1, 2, 3, 3, 9, 2)])
@pytest.mark.parametrize("fwd", [True, False])
def test_check64_costvol_equals_numpy_oracle(B, Cn, h, w, win, F, fwd):
    from oracle import check64 as c64
    r = np.random.default_rng(5)
    fr = [r.standard_normal((B, Cn, h, w)).astype(np.float32) for _ in range(F)]
    wide = r.standard_normal((B, 2 * win * win, h, w)).astype(np.float32)
    np.testing.assert_allclose(c64.costvol_forward(fr, win, fwd), o.costvol_forward(fr, win, fwd), rtol=0, atol=1e-13)
    go = wide[:, win * win:]                                  # batch-strided view, as in the model
    for a, b in zip(c64.costvol_backward(fr, go, win, fwd), o.costvol_backward(fr, go, win, fwd)):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-13)


@pytest.mark.parametrize("B,H,W,Cn,Hg,Wg,sigma", [(2, 9, 11, 3, 9, 11, 4.0), (1, 6, 7, 5, 4, 5, 0.5), (2, 5, 8, 32, 5, 8, 9.0)])
def test_check64_sampler_equals_numpy_oracle(B, H, W, Cn, Hg, Wg, sigma):
    from oracle import check64 as c64
    r = np.random.default_rng(6)
    img = r.standard_normal((B, H, W, Cn)).astype(np.float32)
    grid = (r.standard_normal((B, Hg, Wg, 2)) * sigma).astype(np.float32)
    grid[0, 0, 0] = (0.0, 0.0)
    grid[0, -1, -1] = (0.5, 100.0)        # clamped to the border: tap at H reads 0 (Q3)
    go = r.standard_normal((B, Hg, Wg, Cn)).astype(np.float32)
    np.testing.assert_allclose(c64.warp_forward(img, grid), o.warp_forward(img, grid), rtol=0, atol=1e-13)
    gi, gg = c64.warp_backward(img, grid, go)
    ogi, ogg = o.warp_backward(img, grid, go)
    np.testing.assert_allclose(gi, ogi, rtol=0, atol=1e-12)
    np.testing.assert_allclose(gg, ogg, rtol=0, atol=1e-12)
    assert c64.warp_backward(img, grid, go, only_grid=True)[0] is None


def test_check64_rel_err_matches_the_python_definition_and_flags_nan():
    from oracle import check64 as c64
    r = np.random.default_rng(7)
    b = r.standard_normal((3, 10, 4, 6))
    wide = np.zeros((3, 20, 4, 6), np.float32)
    wide[:, 10:] = b.astype(np.float32)
    wide[1, 13, 2, 3] += 0.25
    a = wide[:, 10:]                                          # strided view (one half of a joined buffer)
    assert abs(c64.rel_err(a, b) - o.rel_err(a, b)) < 1e-12
    wide[2, 15, 0, 0] = np.nan
    assert c64.rel_err(wide[:, 10:], b) == float("inf")
