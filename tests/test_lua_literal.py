"""The closed-form oracle (oracle/b2f_oracle.py: what the CUDA kernels are compared with) against the
statement-by-statement restatements of the reference's Lua (oracle/lua_literal.py on the Torch7 tensor model of
oracle/th7.py).  Two independent readings of criterions/*.lua have to agree to rounding -- in float64 at 1e-12, and
in float32 (the reference's arithmetic type, every temporary rounded where the Lua rounds it) at 1e-5 with identical
out-of-image masks."""
import numpy as np
import pytest

from oracle import b2f_oracle as o
from oracle import lua_literal as lit
from oracle import th7


def _rng(seed=2):   # opts.lua:27: -manualSeed defaults to 2
    return np.random.default_rng(seed)


PENALTIES = {
    "quadratic": (o.QuadraticPenalty, lit.QuadraticPenalty),
    "L1": (o.L1Penalty, lit.L1Penalty),
    "lorentzian": (o.LorentzianPenalty, lit.LorentzianPenalty),
}


def _ob_inputs(r, B, C, h, w, past_flow, flow_sigma):
    flow = (r.standard_normal((B, 2, h, w)) * flow_sigma)
    bflow = (r.standard_normal((B, 2, h, w)) * flow_sigma) if past_flow else None
    e = np.exp(r.standard_normal((B, 2, h, w)))
    occ = e / e.sum(1, keepdims=True)
    w1, w2, tgt = (r.uniform(-2.1, 2.6, (B, C, h, w)) for _ in range(3))
    return flow, bflow, occ, [w1, w2], tgt


def _lua_table(flow, bflow, occ, warped):
    items = [th7.tensor(flow)] + ([th7.tensor(bflow)] if bflow is not None else []) + [th7.tensor(occ)] + \
            [th7.tensor(x) for x in warped]
    return th7.LuaTable(items)


@pytest.mark.parametrize("gradient_terms", [False, True])
@pytest.mark.parametrize("pen", sorted(PENALTIES))
@pytest.mark.parametrize("past_flow", [False, True])
@pytest.mark.parametrize("size_average", [True, False])
def test_ob_criterions_closed_form_equals_the_lua_statements(gradient_terms, pen, past_flow, size_average):
    """OBCCriterion.lua:36-240 / OBGCCriterion.lua:39-300, float64.  Flow large enough (in pixels, after the x20
    scaling) that a good part of the pixels leaves the image: masks, penalties and Q7 are exercised."""
    r = _rng(3)
    B, C, h, w = 2, 3, 6, 7
    flow, bflow, occ, warped, tgt = _ob_inputs(r, B, C, h, w, past_flow, 0.12)
    alpha, beta, gamma = (0.7, 1.3, 0.9) if gradient_terms else (1.0, 1.0, 1.0)
    closed = o.OBCriterionOracle(gradient_terms, PENALTIES[pen][0](), F=3, past_flow=past_flow, pwc_flow_scaling=20,
                                 penalty_out=0.6, size_average=size_average, alpha=alpha, beta=beta, gamma=gamma)
    crit = lit.OBGCCriterion() if gradient_terms else lit.OBCCriterion()
    crit.p, crit.past_flow, crit.pwc_flow_scaling = PENALTIES[pen][1](), past_flow, 20
    crit.penalty_out, crit.sizeAverage = 0.6, size_average
    if gradient_terms:
        crit.alpha, crit.beta, crit.gamma = alpha, beta, gamma
    # the closed form evaluates the out-of-image test in float32 whatever the arithmetic type (Q14): feed both
    # readings float32-representable flows so that the float64 statements see the same coordinates
    flow = flow.astype(np.float32).astype(np.float64)
    bflow = None if bflow is None else bflow.astype(np.float32).astype(np.float64)
    inp = _lua_table(flow, bflow, occ, warped)
    loss = crit.updateOutput(inp, th7.tensor(tgt))
    grads = crit.updateGradInput(inp, th7.tensor(tgt))
    want_loss = closed.forward(flow, bflow, occ, warped, tgt)
    g_occ, g_warp = closed.backward(flow, bflow, occ, warped, tgt)
    assert abs(loss - want_loss) <= 1e-12 * abs(want_loss)
    assert np.allclose(grads[1].a, g_occ, rtol=0, atol=1e-12 * np.abs(g_occ).max())
    for f in range(2):
        assert np.allclose(grads[2 + f].a, g_warp[f], rtol=0, atol=1e-12 * max(np.abs(g_warp[f]).max(), 1e-30))
    # at least one pixel of each warped frame is out of the image and at least one inside: the test means something
    masks = closed._masks(flow, bflow)
    assert all(0 < m.sum() < m.size for m in masks)


@pytest.mark.parametrize("gradient_terms", [False, True])
def test_ob_criterions_float32_statement_order(gradient_terms):
    """The same comparison in the reference's arithmetic type: every Lua temporary is rounded to float32 where Torch7
    rounds it (th7.dtype(float32)); the closed form's float32 mode must stay within 1e-5 and produce the same masks
    (Q14: tcoord = fl(coord + fl(fl(k * flow) * scale)))."""
    r = _rng(4)
    B, C, h, w = 2, 3, 9, 11
    flow, bflow, occ, warped, tgt = _ob_inputs(r, B, C, h, w, True, 0.2)
    f32 = lambda a: None if a is None else a.astype(np.float32)
    flow, bflow, occ, tgt = f32(flow), f32(bflow), f32(occ), f32(tgt)
    warped = [f32(x) for x in warped]
    closed = o.OBCriterionOracle(gradient_terms, o.L1Penalty(), F=3, past_flow=True, pwc_flow_scaling=20,
                                 size_average=False, alpha=0.0 if gradient_terms else 1.0, dtype=np.float32)
    with th7.dtype(np.float32):
        crit = lit.OBGCCriterion() if gradient_terms else lit.OBCCriterion()
        crit.p, crit.past_flow, crit.pwc_flow_scaling, crit.sizeAverage = lit.L1Penalty(), True, 20, False
        if gradient_terms:
            crit.alpha = 0.0            # the README's soft configuration: alpha 0, beta 1, gamma 1
        inp = _lua_table(flow, bflow, occ, warped)
        loss = crit.updateOutput(inp, th7.tensor(tgt))
        grads = crit.updateGradInput(inp, th7.tensor(tgt))
        assert grads[1].a.dtype == np.float32
    want_loss = closed.forward(flow, bflow, occ, warped, tgt)
    g_occ, g_warp = closed.backward(flow, bflow, occ, warped, tgt)
    assert abs(loss - want_loss) <= 1e-5 * abs(want_loss)
    assert o.rel_err(grads[1].a, g_occ) < 1e-5
    for f in range(2):
        assert o.rel_err(grads[2 + f].a, g_warp[f]) < 1e-5
        # hard zeros (masked / out-of-image pixels) are the same set
        assert np.array_equal(grads[2 + f].a == 0, g_warp[f] == 0)


@pytest.mark.parametrize("pen", sorted(PENALTIES))
@pytest.mark.parametrize("size_average", [True, False])
@pytest.mark.parametrize("Cin", [1, 2])
def test_second_order_smoothness_closed_form_equals_the_lua_statements(pen, size_average, Cin):
    """SecondOrderSmoothnessCriterion.lua:28-104 (one backward per forward: :87-88 overwrite self.gy / gx, Q10)."""
    r = _rng(5)
    B, h, w = 2, 7, 8
    x = r.standard_normal((B, Cin, h, w)) * 0.3
    tgt = r.uniform(-2.1, 2.6, (B, 3, h, w))
    closed = o.SmoothnessOracle(2, PENALTIES[pen][0](), cs=20.0, size_average=size_average)
    crit = lit.SecondOrderSmoothnessCriterion()
    crit.p, crit.sizeAverage = PENALTIES[pen][1](), size_average
    loss = crit.updateOutput(th7.tensor(x), th7.tensor(tgt))
    grad = crit.updateGradInput(th7.tensor(x), th7.tensor(tgt))
    want = closed.forward(x, tgt)
    assert abs(loss - want) <= 1e-12 * abs(want)
    g = closed.backward(x, tgt)
    assert np.allclose(grad.a, g, rtol=0, atol=1e-12 * np.abs(g).max())


@pytest.mark.parametrize("pen", sorted(PENALTIES))
@pytest.mark.parametrize("size_average", [True, False])
@pytest.mark.parametrize("Cin", [2, 3])
def test_first_order_smoothness_closed_form_equals_the_lua_statements(pen, size_average, Cin):
    """SmoothnessCriterion.lua:28-106.  Cin = 2 is the model's case (flow / occlusion against the 3-channel image):
    the edge weights then read the re-laid-out buffer of SURVEY Q9 (statements :55-56 replayed on the TH storage
    model); Cin = 3 has no size mismatch and runs entirely on the plain tensor model."""
    r = _rng(9)
    B, h, w = 3, 7, 8
    x = r.standard_normal((B, Cin, h, w)) * 0.3
    tgt = r.uniform(-2.1, 2.6, (B, 3, h, w))
    closed = o.SmoothnessOracle(1, PENALTIES[pen][0](), cs=20.0, size_average=size_average, alias=True)
    crit = lit.SmoothnessCriterion()
    crit.p, crit.sizeAverage = PENALTIES[pen][1](), size_average
    loss = crit.updateOutput(th7.tensor(x), th7.tensor(tgt))
    grad = crit.updateGradInput(th7.tensor(x), th7.tensor(tgt))
    want = closed.forward(x, tgt)
    assert abs(loss - want) <= 1e-12 * abs(want)
    g = closed.backward(x, tgt)
    assert np.allclose(grad.a, g, rtol=0, atol=1e-12 * np.abs(g).max())


@pytest.mark.parametrize("size_average", [True, False])
def test_constvel_closed_form_equals_the_lua_statements(size_average):
    """ConstVelCriterion.lua:29-74 (forward and backward normalisers differ by the channel count, Q11)."""
    r = _rng(6)
    f, b = r.standard_normal((2, 2, 5, 6)), r.standard_normal((2, 2, 5, 6))
    b[0, :, 2, 3] = f[0, :, 2, 3]          # a pixel with zero end-point error: 0 / (0 + 1e-12)
    crit = lit.ConstVelCriterion()
    crit.sizeAverage = size_average
    inp = th7.LuaTable([th7.tensor(f), th7.tensor(b)])
    loss = crit.updateOutput(inp)
    grads = crit.updateGradInput(inp)
    want = o.constvel_forward(f, b, size_average)
    g1, g2 = o.constvel_backward(f, b, size_average)
    assert abs(loss - want) <= 1e-12 * abs(want)
    assert np.allclose(grads[1].a, g1, rtol=0, atol=1e-12) and np.allclose(grads[2].a, g2, rtol=0, atol=1e-12)


@pytest.mark.parametrize("size_average", [True, False])
@pytest.mark.parametrize("C", [2, 3])
def test_occlusion_prior_closed_form_equals_the_lua_statements(size_average, C):
    """OcclusionPriorCriterion.lua:28-73, both channel-count branches (the 'gradient' of :59-66 is not the derivative
    of the forward, Q15)."""
    r = _rng(7)
    e = np.exp(r.standard_normal((2, C, 5, 6)))
    occ = e / e.sum(1, keepdims=True)
    crit = lit.OcclusionPriorCriterion()
    crit.sizeAverage, crit.penalty = size_average, 0.8
    loss = crit.updateOutput(th7.tensor(occ), th7.tensor(occ))
    grad = crit.updateGradInput(th7.tensor(occ), th7.tensor(occ))
    want = o.occprior_forward(occ, size_average, 0.8)
    assert abs(loss - want) <= 1e-12 * abs(want)
    assert np.allclose(grad.a, o.occprior_backward(occ, size_average, 0.8), rtol=0, atol=1e-13)


@pytest.mark.parametrize("win,F,B,Cn,h,w", [(9, 2, 2, 5, 11, 13), (5, 3, 1, 4, 12, 15), (3, 4, 2, 3, 8, 9)])
@pytest.mark.parametrize("fwd", [True, False])
def test_costvol_closed_form_equals_the_lua_statements(win, F, B, Cn, h, w, fwd):
    """models/CostVolMulti.lua:49-181, forward and backward, F = 2..4 frames (the displacement of frame f is
    multiplied by f - 1, :68-69), gradOutput a channel narrow of the joined 2 * win^2 buffer as in pwc.lua:267."""
    r = _rng(8)
    frames = [r.standard_normal((B, Cn, h, w)) for _ in range(F)]
    wide = r.standard_normal((B, 2 * win * win, h, w))
    sl = slice(0, win * win) if fwd else slice(win * win, 2 * win * win)
    m = lit.CostVolMulti(win, fwd)
    inp = th7.LuaTable([th7.tensor(f) for f in frames])
    out = m.updateOutput(inp)
    go = th7.tensor(wide)[th7.ALL, (sl.start + 1, sl.stop), th7.ALL, th7.ALL]      # gradOutput:narrow(2, ...)
    grads = m.updateGradInput(inp, go)
    want = o.costvol_forward(frames, win, fwd)
    assert np.allclose(out.a, want, rtol=0, atol=1e-12 * np.abs(want).max())
    wg = o.costvol_backward(frames, wide[:, sl], win, fwd)
    for f in range(F):
        assert np.allclose(grads[f + 1].a, wg[f], rtol=0, atol=1e-12 * np.abs(wg[f]).max())


def test_tensor_model_views_share_storage_and_overwrite_add():
    """The Torch7 behaviours the restatements lean on."""
    t = th7.tensor(np.arange(24.0).reshape(1, 2, 3, 4))
    v = t[th7.ALL, (2,), th7.ALL, (2, 3)]                  # t[{{},{2},{},{2,3}}]: a view, dimension kept
    assert v.size() == (1, 1, 3, 2)
    v.add(100)                                             # r:add(number) is in place and writes through
    assert t.a[0, 1, 0, 1] == 113 and t.a[0, 0, 0, 1] == 1
    r = th7.Tensor_(1, 1, 3, 2).zero()
    r.add(v, -1, v)                                        # r:add(a, v, b) OVERWRITES with a + v * b
    assert (r.a == 0).all()
    r.fill(5).add(2, v)                                    # r:add(v, b) accumulates v * b
    assert r.a[0, 0, 0, 0] == 5 + 2 * 113
    s = th7.sum_(t, 2)                                     # torch.sum(t, 2) keeps the dimension
    assert s.size() == (1, 1, 3, 4)
    m = th7.ge(t, 12)
    m.cmul(th7.le(t, 13))                                  # Byte masks multiply as AND
    assert m.a.dtype == np.uint8 and m.a.sum() == 0 + (t.a == 12).sum() + (t.a == 13).sum()
    assert (1 - m.cuda()).a.dtype == np.float64            # :cuda() makes the mask a float tensor
    with pytest.raises(NotImplementedError):               # the Q9 resize is not modelled here, loudly
        t[th7.ALL, (1,), th7.ALL, th7.ALL].add(th7.tensor(np.zeros((1, 2, 3, 3))), -1, th7.tensor(np.zeros((1, 2, 3, 3))))
