"""Backward of the conv trunk and of the whole network + Adam on the GPU through the C ABI, against
torch CPU float64 autograd (oracle/pwc_torch.py)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _p(t, off=0):
    return C.c_void_p(t.data_ptr() + 4 * off)


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _smooth(shape, rng, cell=8, lo=-2.1, hi=2.6):
    B, Cn, H, W = shape
    base = rng.uniform(lo, hi, (B, Cn, H // cell + 2, W // cell + 2))
    ys, xs = np.arange(H) / cell, np.arange(W) / cell
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[None, None, :, None], (xs - x0)[None, None, None, :]
    g = lambda dy, dx: base[:, :, y0 + dy][:, :, :, x0 + dx]
    out = (1 - fy) * ((1 - fx) * g(0, 0) + fx * g(0, 1)) + fy * ((1 - fx) * g(1, 0) + fx * g(1, 1))
    return out.astype(np.float32)


CASES = [
    # B, Cin, H, W, Cout, stride
    (2, 16, 16, 32, 32, 1),
    (1, 196, 12, 40, 128, 1),
    (1, 128, 8, 16, 96, 1),
    (1, 32, 9, 36, 2, 1),
    (3, 3, 20, 24, 16, 2),
    (1, 64, 16, 32, 96, 2),
    (1, 5, 10, 38, 7, 1),       # widths the TMA path cannot take
    (1, 6, 5, 19, 4, 2),
    (2, 354, 7, 16, 128, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_conv3x3_backward(case):
    from back2future_b200 import _lib
    from oracle import b2f_oracle as o
    lib = _lib.load()
    B, Cin, H, W, Cout, stride = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((B, Cin, H, W)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, 3, 3)) / np.sqrt(9 * Cin)).astype(np.float32)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    go = rng.standard_normal((B, Cout, Ho, Wo)).astype(np.float32)
    act = rng.standard_normal((B, Cin, H, W)).astype(np.float32)
    xt = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
    wt = torch.from_numpy(w.astype(np.float64)).requires_grad_(True)
    bt = torch.zeros(Cout, dtype=torch.float64, requires_grad=True)
    out = F.conv2d(xt, wt, bt, stride=stride, padding=1)
    out.backward(torch.from_numpy(go.astype(np.float64)))
    ref_gx = xt.grad.numpy() * np.where(act > 0, 1.0, 0.2)
    # device
    wp = torch.empty(int(lib.b2f_conv3x3_packed_floats(Cin, Cout)), device="cuda")
    wd = _dev(w)
    _lib.check(lib.b2f_conv3x3_pack_weights(_p(wd), _p(wp), Cout, Cin, 0, _st()))
    wtp = torch.empty(int(lib.b2f_conv3x3_packed_floats(Cout, Cin)), device="cuda")
    _lib.check(lib.b2f_conv3x3_transpose_packed(_p(wp), _p(wtp), Cout, Cin, _st()))
    god, actd = _dev(go), _dev(act)
    # gin embedded in a wider buffer, accumulate on top of a known value
    wide = torch.full((B, Cin + 3, H, W), 2.0, device="cuda")
    _lib.check(lib.b2f_conv3x3_backward_data(_p(god), 0, _p(wtp), _p(actd), 0, _p(wide, 3 * H * W), (Cin + 3) * H * W, 1,
                                             B, Cin, H, W, Cout, stride, 0.2, _st()))
    torch.cuda.synchronize()
    assert bool((wide[:, :3] == 2.0).all())
    assert o.rel_err(wide[:, 3:].cpu().numpy() - 2.0, ref_gx) < TOL
    gin = torch.empty(B, Cin, H, W, device="cuda")
    _lib.check(lib.b2f_conv3x3_backward_data(_p(god), 0, _p(wtp), None, 0, _p(gin), 0, 0, B, Cin, H, W, Cout, stride, 0.2,
                                             _st()))
    torch.cuda.synchronize()
    assert o.rel_err(gin.cpu().numpy(), xt.grad.numpy()) < TOL
    # weights + bias, accumulated over two calls
    gw = torch.zeros_like(wp)
    gb = torch.zeros(Cout, device="cuda")
    xd = _dev(x)
    for _ in range(2):
        _lib.check(lib.b2f_conv3x3_backward_weights(_p(xd), 0, _p(god), 0, _p(gw), _p(gb), B, Cin, H, W, Cout, stride, _st()))
    gwt = torch.empty(Cout, Cin, 3, 3, device="cuda")
    _lib.check(lib.b2f_conv3x3_pack_weights(_p(gwt), _p(gw), Cout, Cin, 1, _st()))
    torch.cuda.synchronize()
    assert o.rel_err(gwt.cpu().numpy(), 2 * wt.grad.numpy()) < TOL
    assert o.rel_err(gb.cpu().numpy(), 2 * bt.grad.numpy()) < TOL


@pytest.mark.parametrize("shape", [(2, 2, 7, 16), (1, 2, 5, 19), (1, 3, 1, 1), (2, 2, 28, 64)])
def test_small_backward_ops(shape):
    from back2future_b200 import _lib
    from oracle import b2f_oracle as o
    lib = _lib.load()
    rng = np.random.default_rng(3)
    B, Cn, H, W = shape
    x = rng.standard_normal(shape)
    xt = torch.from_numpy(x).requires_grad_(True)
    # bilinear x2
    go = rng.standard_normal((B, Cn, 2 * H, 2 * W))
    F.interpolate(xt, scale_factor=2, mode="bilinear", align_corners=True).backward(torch.from_numpy(go))
    gi = torch.full(shape, 1.0, device="cuda")
    god = _dev(go)      # keep every device operand alive until the synchronize (a temporary's memory is reused at once)
    _lib.check(lib.b2f_upsample_bilinear2x_backward(_p(god), _p(gi), B, Cn, H, W, 1.0, 1, _st()))
    torch.cuda.synchronize()
    assert o.rel_err(gi.cpu().numpy() - 1.0, xt.grad.numpy()) < 2e-5
    # nearest x4
    xt.grad = None
    go4 = rng.standard_normal((B, Cn, 4 * H, 4 * W))
    F.interpolate(xt, scale_factor=4, mode="nearest").backward(torch.from_numpy(go4))
    gi = torch.empty(shape, device="cuda")
    go4d = _dev(go4)
    _lib.check(lib.b2f_upsample_nearest_backward(_p(go4d), _p(gi), B, Cn, H, W, 4, _st()))
    torch.cuda.synchronize()
    assert o.rel_err(gi.cpu().numpy(), xt.grad.numpy()) < 1e-5
    # softmax
    xt.grad = None
    sm = F.softmax(xt, dim=1)
    gs = rng.standard_normal(shape)
    sm.backward(torch.from_numpy(gs))
    gi = torch.empty(shape, device="cuda")
    smd, gsd = _dev(sm.detach().numpy()), _dev(gs)
    _lib.check(lib.b2f_softmax_channels_backward(_p(smd), _p(gsd), _p(gi), B, Cn, H, W, _st()))
    torch.cuda.synchronize()
    assert o.rel_err(gi.cpu().numpy(), xt.grad.numpy()) < 1e-5
    # leaky backward + axpy on strided rows
    g = _dev(gs)
    a = _dev(x)
    _lib.check(lib.b2f_leaky_relu_backward(_p(g), g.numel(), _p(a), a.numel(), g.numel(), 1, 0.2, _st()))
    torch.cuda.synchronize()
    assert o.rel_err(g.cpu().numpy(), gs * np.where(x > 0, 1.0, 0.2)) < 1e-6
    wide = torch.zeros(B, Cn + 2, H, W, device="cuda")
    _lib.check(lib.b2f_axpy2d(_p(wide, 2 * H * W), (Cn + 2) * H * W, _p(a), Cn * H * W, Cn * H * W, B, 0.5, _st()))
    _lib.check(lib.b2f_axpy2d(_p(wide, 2 * H * W), (Cn + 2) * H * W, _p(a), Cn * H * W, Cn * H * W, B, 0.25, _st()))
    torch.cuda.synchronize()
    assert o.rel_err(wide[:, 2:].cpu().numpy(), 0.75 * x) < 1e-6 and bool((wide[:, :2] == 0).all())


def test_adam_step_matches_optim_adam():
    """torch/optim adam.lua restated in numpy float64, three steps."""
    from back2future_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    n = 5000
    x = rng.standard_normal(n).astype(np.float32)
    xd, m, v = _dev(x), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    xr, mr, vr = x.astype(np.float64), np.zeros(n), np.zeros(n)
    lr, b1, b2, eps = 1e-2, 0.9, 0.999, 1e-8
    for t in range(1, 4):
        g = rng.standard_normal(n).astype(np.float32)
        gd = _dev(g)
        _lib.check(lib.b2f_adam_step(_p(xd), _p(gd), _p(m), _p(v), n, lr, b1, b2, eps, 0.0, t, _st()))
        torch.cuda.synchronize()
        mr = b1 * mr + (1 - b1) * g
        vr = b2 * vr + (1 - b2) * g.astype(np.float64) ** 2
        step = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        xr = xr - step * mr / (np.sqrt(vr) + eps)
    torch.cuda.synchronize()
    assert np.abs(xd.cpu().numpy() - xr).max() < 1e-5
    assert lib.b2f_adam_step(_p(xd), _p(xd), _p(m), _p(v), n, lr, b1, b2, eps, 0.0, 0, _st()) != 0


@pytest.mark.parametrize("past_flow,B,H,W", [(False, 1, 64, 128), (True, 1, 64, 64), (False, 2, 64, 64)])
def test_network_backward_matches_autograd(past_flow, B, H, W):
    """model:backward(x, gradOutputs) for random gradOutputs vs autograd of the torch float64 graph: every weight and
    bias gradient of the 62 (Hard) / 92 (Soft) convolutions at 1e-4 (relative to the tensor's scale)."""
    from back2future_b200 import pwc
    from oracle import b2f_oracle as o, pwc_oracle as po, pwc_torch as pt
    oopt = po.Opt(past_flow=past_flow)
    params = po.init_params(oopt, seed=31, scale=2.0)
    net = pwc.PWCNet(pwc.Opt(past_flow=past_flow), params)
    rng = np.random.default_rng(32)
    # smooth textures (what images are): the sampler's flow gradient is piecewise constant in the coordinate, so with
    # white-noise images a pixel whose coordinate sits within fp32 rounding of an integer flips a large term
    x = _smooth((B, 9, H, W), rng)
    xd = _dev(x)
    out = net.forward(xd, graph=False)
    gos = [rng.standard_normal(tuple(t.shape)).astype(np.float32) for t in out]
    net.backward(xd, [_dev(g) for g in gos])
    torch.cuda.synchronize()
    got = net.grad_params()
    ref_out, ref = pt.backward(params, x, gos, oopt)
    for a, b in zip(out, ref_out):
        assert o.rel_err(a.cpu().numpy(), b) < TOL
    worst = ("", 0.0)
    for k in ref:
        e = o.rel_err(got[k], ref[k])
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < TOL, worst
    # a second backward (graph replay) reproduces it: gradients are zeroed, not accumulated across steps
    net.backward(xd, [_dev(g) for g in gos], graph=True)
    net.backward(xd, [_dev(g) for g in gos], graph=True)
    torch.cuda.synchronize()
    again = net.grad_params()
    for k in ref:
        assert o.rel_err(again[k], ref[k]) < TOL, k
    # state_params round-trips the weights
    sp = net.state_params()
    for k in params:
        assert np.array_equal(sp[k], params[k]), k
