"""The conv-trunk oracle (oracle/pwc_oracle.py) pinned against torch's CPU float64 functional ops -- the descendants
of the THNN routines Torch7 calls.  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pwc_oracle as po


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("shape", [(2, 3, 8, 12), (1, 5, 7, 9), (1, 16, 1, 4)])
def test_conv3x3_matches_torch(stride, shape):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape)
    w = rng.standard_normal((6, shape[1], 3, 3))
    b = rng.standard_normal(6)
    ref = F.conv2d(_t(x), _t(w), _t(b), stride=stride, padding=1).numpy()
    got = po.conv3x3(x, w, b, stride)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("shape", [(2, 2, 3, 5), (1, 3, 1, 1), (1, 1, 7, 16), (1, 2, 2, 2)])
def test_upsample_bilinear2x_is_align_corners(shape):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape)
    ref = F.interpolate(_t(x), scale_factor=2, mode="bilinear", align_corners=True).numpy()
    np.testing.assert_allclose(po.upsample_bilinear2x(x), ref, rtol=1e-12, atol=1e-12)


def test_small_modules_match_torch():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 3, 6, 10))
    np.testing.assert_allclose(po.avgpool2x2(x), F.avg_pool2d(_t(x), 2, 2).numpy(), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(po.avgpool2x2(x[:, :, :5, :7]), F.avg_pool2d(_t(x[:, :, :5, :7]), 2, 2).numpy(),
                               rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(po.upsample_nearest(x, 2), F.interpolate(_t(x), scale_factor=2, mode="nearest").numpy())
    np.testing.assert_allclose(po.spatial_softmax(x), F.softmax(_t(x), dim=1).numpy(), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(po.leaky_relu(x), F.leaky_relu(_t(x), 0.2).numpy())


def test_parameter_counts_match_the_survey():
    """7 193 316 parameters (hard), 10 168 302 with the past-flow decoders."""
    assert po.n_params(po.Opt()) == 7193316
    assert po.n_params(po.Opt(past_flow=True)) == 10168302


@pytest.mark.parametrize("past_flow", [False, True])
def test_network_output_table_shapes(past_flow):
    """pwc.lua:459-489 / SURVEY App. B: 5 levels at (H, W) / 2^{0..4}; {flow, [bflow,] occ, warp1, warp3}."""
    opt = po.Opt(past_flow=past_flow)
    params = po.init_params(opt, seed=3)
    rng = np.random.default_rng(4)
    x = rng.uniform(-2.1, 2.6, (1, 9, 64, 128)).astype(np.float32)
    out = po.pwc_forward(params, x, opt)
    per = 5 if past_flow else 4
    assert len(out) == 5 * per
    for i in range(5):
        h, w = 64 >> i, 128 >> i
        unit = out[i * per:(i + 1) * per]
        chans = [2, 2, 2, 3, 3] if past_flow else [2, 2, 3, 3]
        for t, c in zip(unit, chans):
            assert t.shape == (1, c, h, w)
        occ = unit[2 if past_flow else 1]
        np.testing.assert_allclose(occ.sum(axis=1), 1.0, rtol=1e-12)
    assert po.flow_scales(opt) == [1.25, 2.5, 5.0, 10.0, 20.0]


def test_network_against_a_torch_functional_composition():
    """The wiring (joins, scales, level order) restated a second time with torch ops + the sampler oracle."""
    from oracle import b2f_oracle as o
    opt = po.Opt()
    params = po.init_params(opt, seed=5, scale=2.0)
    rng = np.random.default_rng(6)
    x = rng.uniform(-2.1, 2.6, (1, 9, 64, 64)).astype(np.float32)
    P = {k: _t(v) for k, v in params.items()}

    def unit(name, t):
        t = F.leaky_relu(F.conv2d(t, P[name + ".0.weight"], P[name + ".0.bias"], stride=2, padding=1), 0.2)
        return F.leaky_relu(F.conv2d(t, P[name + ".1.weight"], P[name + ".1.bias"], padding=1), 0.2)

    def dec(name, t):
        for i in range(6):
            t = F.conv2d(t, P["%s.%d.weight" % (name, i)], P["%s.%d.bias" % (name, i)], padding=1)
            if i < 5:
                t = F.leaky_relu(t, 0.2)
        return t

    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    xt = _t(x)
    I = [xt[:, 0:3], xt[:, 3:6], xt[:, 6:9]]
    cs = []
    for f in range(3):
        pyr = [I[f]]
        for l in range(2, 8):
            pyr.append(unit("feat.l%d" % l, pyr[-1]))
        cs.append(pyr)
    ds = {0: [I[0]], 2: [I[2]]}
    for f in (0, 2):
        for _ in range(4):
            ds[f].append(F.avg_pool2d(ds[f][-1], 2, 2))
    ufs, warped, outs = None, None, {}
    for l in range(7, 2, -1):
        refl = cs[1][l - 1]
        fut = cs[2][l - 1] if l == 7 else warped[2]
        past = cs[0][l - 1] if l == 7 else warped[0]
        cv = torch.cat([_t(o.costvol_forward([refl.numpy(), fut.numpy()], 9, True)),
                        _t(o.costvol_forward([refl.numpy(), past.numpy()], 9, False))], 1)
        join = [cv, refl] + ([ufs] if l < 7 else [])
        occ = F.softmax(dec("occ.l%d" % l, torch.cat(join, 1)), dim=1)
        flow = dec("flow.l%d" % l, cv if l == 7 else torch.cat(join, 1))
        ufs = up(flow)
        sk = up(ufs)
        outs[l] = [sk, F.interpolate(occ, scale_factor=4, mode="nearest")]
        warped = {}
        for f, sgn in ((0, -1), (2, 1)):
            if l > 3:
                warped[f] = _t(o.warping_unit_forward(cs[f][l - 2].numpy().astype(np.float32),
                                                      ufs.numpy().astype(np.float32), 20 * sgn / 2.0 ** (l - 2)))
            outs[l].append(_t(o.warping_unit_forward(ds[f][l - 3].numpy().astype(np.float32),
                                                     sk.numpy().astype(np.float32), 20 * sgn / 2.0 ** (l - 3))))
    got = po.pwc_forward(params, x, opt)
    k = 0
    for l in range(3, 8):
        for t in outs[l]:
            assert o.rel_err(got[k], t.numpy()) < 1e-9, (l, k)
            k += 1
