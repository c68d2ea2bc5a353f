"""Tensor-core 3x3 convolution (tcgen05, three-pass TF32 split) against the float64 oracle, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
TOL = 1e-4


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _p(t, off=0):
    return C.c_void_p(t.data_ptr() + 4 * off) if t is not None else None


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _run(lib, _lib, x, w, b, slope, want_split, want_planar, x_c_total=None, x_c0=0):
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    cinp, coutp = (Cin + 31) // 32 * 32, (Cout + 31) // 32 * 32
    xt = x_c_total or Cin
    xw = torch.zeros(B, xt, H, W, device="cuda")
    xw[:, x_c0:x_c0 + Cin] = _dev(x)
    xh, xl = torch.empty(B, H, W, cinp, device="cuda"), torch.empty(B, H, W, cinp, device="cuda")
    _lib.check(lib.b2f_nhwc_split_from_bdhw(_p(xw, x_c0 * H * W), xt * H * W, _p(xh), _p(xl), B, Cin, H, W, _st()))
    n = int(lib.b2f_conv3x3_tc_packed_floats(Cin, Cout))
    wh, wl = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    wd = _dev(w)
    _lib.check(lib.b2f_conv3x3_tc_pack_weights(_p(wd), _p(wh), _p(wl), Cout, Cin, _st()))
    bd = _dev(b) if b is not None else None
    oh = torch.full((B, H, W, coutp), 9.0, device="cuda") if want_split else None
    ol = torch.full((B, H, W, coutp), 9.0, device="cuda") if want_split else None
    op = torch.full((B, Cout + 2, H, W), 7.0, device="cuda") if want_planar else None
    _lib.check(lib.b2f_conv3x3_tc_forward(_p(xh), _p(xl), _p(wh), _p(wl), _p(bd), _p(oh), _p(ol),
                                          _p(op, 2 * H * W) if want_planar else None, (Cout + 2) * H * W if want_planar else 0,
                                          B, Cin, H, W, Cout, slope, _st()))
    torch.cuda.synchronize()
    # the split of the input is exact and its pad channels are zero
    xs = (xh.double() + xl.double()).cpu().numpy()
    assert np.array_equal(xs[..., :Cin], x.transpose(0, 2, 3, 1).astype(np.float64))
    assert not xs[..., Cin:].any()
    assert bool(((xh.view(torch.int32) & 0x1FFF) == 0).all())
    return oh, ol, op


CASES = [
    # B, Cin, H, W, Cout
    (1, 196, 21, 40, 128),     # decoder conv 0 at level 3: channels padded 196 -> 224, ragged tile columns
    (2, 128, 14, 32, 128),
    (1, 128, 16, 32, 96),
    (1, 96, 9, 20, 64),
    (1, 64, 7, 16, 32),
    (1, 354, 7, 16, 128),      # coarsest level
    (2, 64, 10, 20, 192),      # more than 128 output channels: slices of one launch (blockIdx.z), (hi, lo) outputs too
    (8, 128, 20, 40, 128),     # 72 tiles: 32-column slices
    (8, 96, 40, 80, 96),       # 240 tiles: one slice again
    (1, 32, 30, 50, 64),       # ragged in both directions
]


@pytest.mark.parametrize("case", CASES)
def test_conv3x3_tc_forward(case):
    from back2future_b200 import _lib
    from oracle import b2f_oracle as o, pwc_oracle as po
    lib = _lib.load()
    B, Cin, H, W, Cout = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((B, Cin, H, W)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, 3, 3)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    ref = po.leaky_relu(po.conv3x3(x, w, b, 1))
    oh, ol, op = _run(lib, _lib, x, w, b, 0.2, True, True, x_c_total=Cin + 3, x_c0=3)
    got = (oh.double() + ol.double()).cpu().numpy()
    assert o.rel_err(got[..., :Cout].transpose(0, 3, 1, 2), ref) < TOL
    assert not got[..., Cout:].any()
    assert bool(((oh.view(torch.int32) & 0x1FFF) == 0).all())        # hi is exactly representable in TF32
    assert o.rel_err(op[:, 2:].cpu().numpy(), ref) < TOL
    assert bool((op[:, :2] == 7.0).all())
    # planar only, no bias, no activation
    ref2 = po.conv3x3(x, w, None, 1)
    _, _, op2 = _run(lib, _lib, x, w, None, 1.0, False, True)
    assert o.rel_err(op2[:, 2:].cpu().numpy(), ref2) < TOL


def test_conv3x3_tc_rejects_bad_arguments():
    from back2future_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(1, 8, 8, 32, device="cuda")
    w = torch.zeros(int(lib.b2f_conv3x3_tc_packed_floats(32, 32)), device="cuda")
    o = torch.zeros(1, 8, 8, 32, device="cuda")
    assert lib.b2f_conv3x3_tc_forward(_p(t), _p(t), _p(w), _p(w), None, _p(o), None, None, 0, 1, 32, 8, 8, 32, 0.2, _st()) != 0
    assert lib.b2f_conv3x3_tc_forward(_p(t), _p(t), _p(w), _p(w), None, None, None, None, 0, 1, 32, 8, 8, 32, 0.2, _st()) != 0
    assert lib.b2f_conv3x3_tc_forward(_p(t), _p(t), _p(w), _p(w), None, _p(o), _p(o), None, 0, 1, 32, 0, 8, 32, 0.2, _st()) != 0
    assert b"bad size" in lib.b2f_last_error()
    # the weight gradient's activation tensor must hold at least the convolution's input channels
    assert lib.b2f_conv3x3_tc_backward_weights(_p(t), _p(t), 16, _p(o), _p(o), None, 0, _p(w), None, 1, 32, 8, 8, 32, _st()) != 0


WGRAD_CASES = [
    # B, Cx (channels of the activation tensor), Cin (of the convolution), H, W, Cout
    (2, 128, 128, 14, 32, 128),
    (1, 196, 196, 21, 40, 128),    # decoder conv 0 at level 3: two input-channel chunks (128 + 96 of 224), ragged tile edges
    (1, 128, 128, 16, 32, 96),
    (1, 96, 96, 9, 20, 64),        # M = 128 with one zero block
    (1, 64, 64, 7, 16, 32),
    (1, 354, 162, 7, 16, 128),     # coarsest flow decoder: the first 162 channels of a wider joined input
    (3, 128, 128, 33, 50, 128),    # more tiles than fit one pass of the 3-stage ring, odd sizes
]


@pytest.mark.parametrize("B,Cx,Cin,H,W,Cout", WGRAD_CASES)
def test_conv3x3_tc_backward_weights(B, Cx, Cin, H, W, Cout):
    """SpatialConvolution:accGradParameters on the tensor cores (MN-major operands over the channel-minor (hi, lo)
    tensors, wgrad_tc.cu) against a float64 correlation, accumulating into a gradient buffer that already holds
    values."""
    from back2future_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(7)
    x = rng.standard_normal((B, Cx, H, W)).astype(np.float32)
    g = rng.standard_normal((B, Cout, H, W)).astype(np.float32)
    cxp, coutp32 = (Cx + 31) // 32 * 32, (Cout + 31) // 32 * 32
    xd, gd = _dev(x), _dev(g)
    xh, xl = torch.empty(B, H, W, cxp, device="cuda"), torch.empty(B, H, W, cxp, device="cuda")
    gh, gl = torch.empty(B, H, W, coutp32, device="cuda"), torch.empty(B, H, W, coutp32, device="cuda")
    _lib.check(lib.b2f_nhwc_split_from_bdhw(_p(xd), 0, _p(xh), _p(xl), B, Cx, H, W, _st()))
    _lib.check(lib.b2f_nhwc_split_from_bdhw(_p(gd), 0, _p(gh), _p(gl), B, Cout, H, W, _st()))
    n = int(lib.b2f_conv3x3_packed_floats(Cin, Cout))
    coutp = (Cout + 63) // 64 * 64
    init = rng.standard_normal(n).astype(np.float32)
    gw, gb = _dev(init), torch.full((Cout,), 0.5, device="cuda")
    _lib.check(lib.b2f_conv3x3_tc_backward_weights(_p(xh), _p(xl), Cx, _p(gh), _p(gl), _p(gd), 0, _p(gw), _p(gb), B, Cin, H, W,
                                                   Cout, _st()))
    torch.cuda.synchronize()
    xp = np.pad(x[:, :Cin].astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    want = np.zeros((Cin, 9, coutp))
    g64 = g.astype(np.float64)
    for ky in range(3):
        for kx in range(3):
            want[:, ky * 3 + kx, :Cout] = np.einsum("bchw,bnhw->cn", xp[:, :, ky:ky + H, kx:kx + W], g64)
    got = gw.cpu().numpy().reshape(Cin, 9, coutp).astype(np.float64) - init.reshape(Cin, 9, coutp)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() < 1e-4 * scale, np.abs(got - want).max() / scale
    assert not got[:, :, Cout:].any()
    wb = g64.sum(axis=(0, 2, 3)) + 0.5
    assert np.abs(gb.cpu().numpy() - wb).max() < 1e-4 * np.abs(wb).max()


def test_conv3x3_tc_backward_weights_bias_from_split_gradient():
    """g_planar == NULL: the bias gradient is summed from the channel-minor (hi, lo) output gradient."""
    from back2future_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(8)
    B, Cin, H, W, Cout = 2, 64, 13, 24, 96
    x = rng.standard_normal((B, Cin, H, W)).astype(np.float32)
    g = rng.standard_normal((B, Cout, H, W)).astype(np.float32)
    xd, gd = _dev(x), _dev(g)
    xh, xl = torch.empty(B, H, W, Cin, device="cuda"), torch.empty(B, H, W, Cin, device="cuda")
    gh, gl = torch.empty(B, H, W, Cout, device="cuda"), torch.empty(B, H, W, Cout, device="cuda")
    _lib.check(lib.b2f_nhwc_split_from_bdhw(_p(xd), 0, _p(xh), _p(xl), B, Cin, H, W, _st()))
    _lib.check(lib.b2f_nhwc_split_from_bdhw(_p(gd), 0, _p(gh), _p(gl), B, Cout, H, W, _st()))
    gw = torch.zeros(int(lib.b2f_conv3x3_packed_floats(Cin, Cout)), device="cuda")
    gb = torch.full((Cout,), -0.25, device="cuda")
    _lib.check(lib.b2f_conv3x3_tc_backward_weights(_p(xh), _p(xl), Cin, _p(gh), _p(gl), None, 0, _p(gw), _p(gb), B, Cin, H, W,
                                                   Cout, _st()))
    torch.cuda.synchronize()
    wb = g.astype(np.float64).sum(axis=(0, 2, 3)) - 0.25
    assert np.abs(gb.cpu().numpy() - wb).max() < 1e-4 * np.abs(wb).max()


DGRAD_CASES = [
    # B, Cout (K), H, W, Cin (N), mask
    (2, 128, 14, 32, 128, "hi"),
    (1, 128, 9, 20, 96, "planar"),
    (1, 64, 7, 16, 32, "hi"),
    (1, 96, 16, 33, 64, "none"),
    (1, 128, 10, 20, 196, "slices"),      # first decoder layer: slices of <= 128, planar, second call accumulates
    (3, 192, 5, 10, 192, "slices_hi"),    # coarsest pyramid layer: slices with the channel-minor mask
]


@pytest.mark.parametrize("B,Cout,H,W,Cin,mask", DGRAD_CASES)
def test_conv3x3_tc_backward_data(B, Cout, H, W, Cin, mask):
    """SpatialConvolution:updateGradInput on the tensor cores against the float64 transpose of the oracle's
    convolution: gin[ci, y, x] = sum_{co, ky, kx} g[co, y + 1 - ky, x + 1 - kx] w[co, ci, ky, kx], times the
    LeakyReLU derivative of the layer below taken from its planar output or from the HI half of its split."""
    from back2future_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)
    g = rng.standard_normal((B, Cout, H, W)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, 3, 3)) * 0.05).astype(np.float32)
    act = rng.standard_normal((B, Cin, H, W)).astype(np.float32)
    act[rng.random(act.shape) < 0.05] = 0.0                 # exact zeros take the slope branch (act > 0 is false)
    coutp, cinp = (Cout + 31) // 32 * 32, (Cin + 31) // 32 * 32
    gd, ad = _dev(g), _dev(act)
    gh, gl = torch.empty(B, H, W, coutp, device="cuda"), torch.empty(B, H, W, coutp, device="cuda")
    _lib.check(lib.b2f_nhwc_split_from_bdhw(_p(gd), 0, _p(gh), _p(gl), B, Cout, H, W, _st()))
    ah, al = torch.empty(B, H, W, cinp, device="cuda"), torch.empty(B, H, W, cinp, device="cuda")
    _lib.check(lib.b2f_nhwc_split_from_bdhw(_p(ad), 0, _p(ah), _p(al), B, Cin, H, W, _st()))
    wp = torch.empty(int(lib.b2f_conv3x3_packed_floats(Cin, Cout)), device="cuda")
    _lib.check(lib.b2f_conv3x3_pack_weights(_p(_dev(w)), _p(wp), Cout, Cin, 0, _st()))
    nt = 9 * Cin * coutp
    th, tl = torch.empty(nt, device="cuda"), torch.empty(nt, device="cuda")
    _lib.check(lib.b2f_conv3x3_tc_pack_from_packed(_p(wp), _p(th), _p(tl), Cout, Cin, Cout, 1, _st()))
    # float64 reference
    gp = np.pad(g.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    want = np.zeros((B, Cin, H, W))
    w64 = w.astype(np.float64)
    for ky in range(3):
        for kx in range(3):
            want += np.einsum("bnhw,nc->bchw", gp[:, :, 2 - ky:2 - ky + H, 2 - kx:2 - kx + W], w64[:, :, ky, kx])
    slope = 0.2
    masked = mask in ("hi", "planar", "slices_hi")
    if masked:
        want = np.where(act > 0, want, slope * want)
    sliced = mask.startswith("slices")
    one = Cin in (32, 64, 96, 128)
    oh = torch.full((B, H, W, cinp), 9.0, device="cuda") if one else None
    ol = torch.full((B, H, W, cinp), 9.0, device="cuda") if one else None
    base = rng.standard_normal((B, Cin + 1, H, W)).astype(np.float32)
    op = _dev(base) if (sliced or mask == "planar") else None
    acc = 1 if sliced else 0
    _lib.check(lib.b2f_conv3x3_tc_backward_data(
        _p(gh), _p(gl), _p(th), _p(tl), _p(ad) if mask == "planar" else None, 0, _p(ah) if mask in ("hi", "slices_hi") else None,
        _p(oh), _p(ol), _p(op, H * W) if op is not None else None, (Cin + 1) * H * W if op is not None else 0, B, Cout, H, W,
        Cin, slope, acc, _st()))
    torch.cuda.synchronize()
    scale = np.abs(want).max()
    if oh is not None:
        got = (oh.double() + ol.double()).cpu().numpy()
        assert np.abs(got[..., :Cin].transpose(0, 3, 1, 2) - want).max() < TOL * scale
        assert not got[..., Cin:].any()
    if op is not None:
        got = op.cpu().numpy().astype(np.float64)
        assert np.array_equal(got[:, 0], base[:, 0].astype(np.float64))            # the channel in front is untouched
        ref = want + (base[:, 1:].astype(np.float64) if acc else 0.0)
        assert np.abs(got[:, 1:] - ref).max() < TOL * max(scale, np.abs(ref).max())
    # alternatives are exclusive
    rc = lib.b2f_conv3x3_tc_backward_data(_p(gh), _p(gl), _p(th), _p(tl), _p(ad), 0, _p(ah), _p(oh), _p(ol),
                                          _p(op, H * W) if op is not None else None, 0, B, Cout, H, W, Cin, slope, 0, _st())
    assert rc != 0


S2_CASES = [
    # B, Cout (K), H, W (input-gradient plane), Cin (N), accumulate
    (2, 64, 16, 32, 32, 0),
    (1, 96, 10, 20, 64, 1),
    (1, 32, 9, 15, 16, 1),       # odd sizes, 16 valid channels of a 32-wide slice
    (1, 128, 14, 40, 96, 0),
    (2, 192, 10, 20, 128, 1),    # four 128-column accumulators: all 512 TMEM columns
]


@pytest.mark.parametrize("B,Cout,H,W,Cin,acc", S2_CASES)
def test_conv3x3_tc_backward_data_stride2(B, Cout, H, W, Cin, acc):
    """Input gradient of a stride-2 convolution on the tensor cores (four parity-class accumulators) against the
    float64 scatter gin[ci, 2 yo + ky - 1, 2 xo + kx - 1] += g[co, yo, xo] w[co, ci, ky, kx]."""
    from back2future_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(13)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    g = rng.standard_normal((B, Cout, Ho, Wo)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, 3, 3)) * 0.05).astype(np.float32)
    coutp = (Cout + 31) // 32 * 32
    gd = _dev(g)
    gh, gl = torch.empty(B, Ho, Wo, coutp, device="cuda"), torch.empty(B, Ho, Wo, coutp, device="cuda")
    _lib.check(lib.b2f_nhwc_split_from_bdhw(_p(gd), 0, _p(gh), _p(gl), B, Cout, Ho, Wo, _st()))
    wp = torch.empty(int(lib.b2f_conv3x3_packed_floats(Cin, Cout)), device="cuda")
    _lib.check(lib.b2f_conv3x3_pack_weights(_p(_dev(w)), _p(wp), Cout, Cin, 0, _st()))
    nt = 9 * Cin * coutp
    th, tl = torch.empty(nt, device="cuda"), torch.empty(nt, device="cuda")
    _lib.check(lib.b2f_conv3x3_tc_pack_from_packed(_p(wp), _p(th), _p(tl), Cout, Cin, Cout, 1, _st()))
    want = np.zeros((B, Cin, H + 2, W + 2))
    g64, w64 = g.astype(np.float64), w.astype(np.float64)
    for ky in range(3):
        for kx in range(3):
            # padded coordinates: Y + 1 = 2 yo + ky
            want[:, :, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2] += np.einsum("bnhw,nc->bchw", g64, w64[:, :, ky, kx])
    want = want[:, :, 1:H + 1, 1:W + 1]
    base = rng.standard_normal((B, Cin + 1, H, W)).astype(np.float32)
    op = _dev(base)
    _lib.check(lib.b2f_conv3x3_tc_backward_data_s2(_p(gh), _p(gl), _p(th), _p(tl), _p(op, H * W), (Cin + 1) * H * W, B, Cout, Ho,
                                                   Wo, Cin, H, W, acc, _st()))
    torch.cuda.synchronize()
    got = op.cpu().numpy().astype(np.float64)
    assert np.array_equal(got[:, 0], base[:, 0].astype(np.float64))
    ref = want + (base[:, 1:].astype(np.float64) if acc else 0.0)
    assert np.abs(got[:, 1:] - ref).max() < TOL * max(np.abs(want).max(), np.abs(ref).max())
    assert lib.b2f_conv3x3_tc_backward_data_s2(_p(gh), _p(gl), _p(th), _p(tl), _p(op), 0, B, Cout, Ho + 1, Wo, Cin, H, W, 0,
                                               _st()) != 0


def test_conv3x3_tc_pack_from_packed_batch_matches_the_single_calls():
    """One launch over a device table of b2f_pack_job entries against the per-tensor entry point, forward and transposed."""
    from back2future_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(21)
    shapes = [(128, 196, 224, 0), (2, 32, 32, 0), (96, 128, 128, 0), (128, 196, 128, 1), (2, 32, 2, 1), (64, 32, 64, 1)]
    jobs, singles, keep = [], [], []
    for cout, cin, K, tr in shapes:
        w = _dev((rng.standard_normal((cout, cin, 3, 3)) * 0.1).astype(np.float32))
        wp = torch.empty(int(lib.b2f_conv3x3_packed_floats(cin, cout)), device="cuda")
        _lib.check(lib.b2f_conv3x3_pack_weights(_p(w), _p(wp), cout, cin, 0, _st()))
        n = 9 * (cin if tr else cout) * ((K + 31) // 32 * 32)
        a = [torch.full((n,), 7.0, device="cuda") for _ in range(4)]
        _lib.check(lib.b2f_conv3x3_tc_pack_from_packed(_p(wp), _p(a[0]), _p(a[1]), cout, cin, K, tr, _st()))
        jobs.append((wp.data_ptr(), a[2].data_ptr(), a[3].data_ptr(), cout, cin, K, tr))
        singles.append(a)
        keep += [w, wp]
    table = torch.frombuffer(bytearray(_lib.pack_jobs(jobs)), dtype=torch.uint8).cuda()
    _lib.check(lib.b2f_conv3x3_tc_pack_from_packed_batch(_p(table), len(jobs), _st()))
    torch.cuda.synchronize()
    for a in singles:
        assert torch.equal(a[0], a[2]) and torch.equal(a[1], a[3])
    assert lib.b2f_conv3x3_tc_pack_from_packed_batch(None, 3, _st()) != 0
    assert lib.b2f_conv3x3_tc_pack_from_packed_batch(None, 0, _st()) == 0
