"""CPU tests over the committed golden fixture (tests/golden/hotpath_golden.npz, made by
tests/golden/make_golden.py from crops of the reference's sample frames).

The expected values are oracle-derived (the Lua/Torch7 reference cannot run here; parity is unpinned,
SURVEY 8c) -- these tests pin the numpy oracle against silent change, check the independent C restatement
against the same vectors, and guard the fixture itself.  The GPU side of the same fixture is in
tests/test_gpu_parity.py (test_golden_*).
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import b2f_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz")
TIGHT = 2e-7      # float64 oracle output stored as float32


@pytest.fixture(scope="module")
def G():
    with np.load(PATH) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def cpu_lib():
    path = os.path.join(ROOT, "oracle", "c", "libb2f_cpu.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build_oracle()
    return C.CDLL(path)


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_fixture_is_complete_and_float32(G):
    assert len(G) == 53
    for k, v in G.items():
        assert v.dtype in (np.float32, np.float64, np.bool_), k
        assert np.isfinite(v).all(), k
    assert G["frame_ref"].shape == (2, 3, 20, 40) and G["feat_ref"].shape == (2, 8, 12, 24)
    # the flows do leave the image: both mask values occur, and the sampler clamp is hit
    assert 0 < (~G["mask_fut"]).sum() < G["mask_fut"].size
    x = np.arange(40, dtype=np.float32)[None, None, :] + G["grid_fut"][..., 0]
    assert (x < 0).any() and (x > 39).any()


def test_costvol_oracle_reproduces_golden(G):
    for name, frm, sl, fwd in (("cv_fwd", "feat_fut", slice(0, 81), True), ("cv_bwd", "feat_past", slice(81, 162), False)):
        frames = [G["feat_ref"], G[frm]]
        assert o.rel_err(o.costvol_forward(frames, 9, fwd), G[name + "_out"]) < TIGHT
        assert o.rel_err(o.costvol_forward_lua(frames, 9, fwd), G[name + "_out"]) < TIGHT   # literal Lua loops
        gr, gf = o.costvol_backward(frames, G["cv_gradout_joined"][:, sl], 9, fwd)
        assert o.rel_err(gr, G[name + "_gradref"]) < TIGHT and o.rel_err(gf, G[name + "_gradframe"]) < TIGHT


def test_warp_oracle_reproduces_golden(G):
    img = np.ascontiguousarray(G["frame_fut"].transpose(0, 2, 3, 1))
    assert o.rel_err(o.warp_forward(img, G["grid_fut"]), G["warp_img_fut"]) < TIGHT
    assert o.rel_err(o.warp_forward_loops(img, G["grid_fut"]), G["warp_img_fut"]) < 1e-5   # scalar fp32 restatement
    gi, gg = o.warp_backward(img, G["grid_fut"], G["warp_gradout3"])
    assert o.rel_err(gi, G["warp_img_fut_gradimg"]) < TIGHT and o.rel_err(gg, G["warp_img_fut_gradgrid"]) < TIGHT
    ft = np.ascontiguousarray(G["feat_fut_full"].transpose(0, 2, 3, 1))
    assert o.rel_err(o.warp_forward(ft, G["grid_fut"]), G["warp_feat_fut"]) < TIGHT
    gi, gg = o.warp_backward(ft, G["grid_fut"], G["warp_gradout8"])
    assert o.rel_err(gi, G["warp_feat_fut_gradimg"]) < TIGHT and o.rel_err(gg, G["warp_feat_fut_gradgrid"]) < TIGHT
    # the grids are MulConstant(+-20) of the network-unit flow (pwc.lua:443), BHW2 with x first
    assert np.array_equal(G["grid_fut"], (G["flow"] * np.float32(20)).transpose(0, 2, 3, 1))
    assert np.array_equal(G["grid_past"], (G["flow"] * np.float32(-20)).transpose(0, 2, 3, 1))


def test_criterion_oracle_reproduces_golden(G):
    flow, bflow, occ, ref = G["flow"], G["bflow"], G["occ"], G["frame_ref"]
    wp, wf = G["crit_warp_past"], G["crit_warp_fut"]
    for name, gt, pf, alpha in (("obcc", False, False, 1.0), ("obgcc", True, True, 0.0)):
        oc = o.OBCriterionOracle(gt, o.L1Penalty(), past_flow=pf, pwc_flow_scaling=20.0, size_average=False, alpha=alpha)
        bf = bflow if pf else None
        assert abs(oc.forward(flow, bf, occ, [wp, wf], ref) - G[name + "_loss"]) < 1e-12 * abs(G[name + "_loss"])
        ro, rw = oc.backward(flow, bf, occ, [wp, wf], ref)
        assert o.rel_err(ro, G[name + "_gradocc"]) < TIGHT
        assert o.rel_err(rw[0], G[name + "_gradwarp_past"]) < TIGHT and o.rel_err(rw[1], G[name + "_gradwarp_fut"]) < TIGHT
    for name, order, inp, pen, alias in (("smooth1_flow", 1, flow, 1, True), ("smooth2_flow", 2, flow, 1, True),
                                         ("smooth1_occ", 1, occ, 0, True), ("smooth1_flow_intended", 1, flow, 1, False)):
        oc = o.SmoothnessOracle(order, o.make_penalty(pen), size_average=False, alias=alias)
        assert abs(oc.forward(inp, ref) - G[name + "_loss"]) < 1e-12 * abs(G[name + "_loss"])
        assert o.rel_err(oc.backward(inp, ref), G[name + "_grad"]) < TIGHT
    # Q9: on real images the aliased weights are not the intended ones
    assert abs(G["smooth1_flow_loss"] - G["smooth1_flow_intended_loss"]) > 1e-3 * abs(G["smooth1_flow_loss"])
    assert abs(o.constvel_forward(flow, bflow, True) - G["constvel_loss"]) < 1e-12
    a, b = o.constvel_backward(flow, bflow, True)
    assert o.rel_err(a, G["constvel_gradf"]) < TIGHT and o.rel_err(b, G["constvel_gradb"]) < TIGHT
    assert abs(o.occprior_forward(occ, False) - G["occprior_loss"]) < 1e-9
    assert o.rel_err(o.occprior_backward(occ, False), G["occprior_grad"]) < TIGHT
    assert np.array_equal(o.out_of_image_mask(flow, 1, 20.0), G["mask_fut"])
    assert np.array_equal(o.out_of_image_mask(bflow, -1, 20.0), G["mask_past_bflow"])


def test_c_restatement_reproduces_golden(G, cpu_lib):
    """The C/OpenMP restatement (the CPU baseline of bench.py) on the same real-image vectors, fp32."""
    for name, frm, off, fwd in (("cv_fwd", "feat_fut", 0, 1), ("cv_bwd", "feat_past", 81, 0)):
        frames = [np.ascontiguousarray(G["feat_ref"]), np.ascontiguousarray(G[frm])]
        B, Cn, h, w = frames[0].shape
        ptrs = (C.c_void_p * 2)(*[f.ctypes.data for f in frames])
        out = np.empty((B, 81, h, w), np.float32)
        assert cpu_lib.b2fcpu_costvol_forward(ptrs, 2, B, Cn, h, w, 9, fwd, _fp(out)) == 0
        assert o.rel_err(out, G[name + "_out"]) < 1e-5
        wide = np.ascontiguousarray(G["cv_gradout_joined"])
        grads = [np.empty_like(f) for f in frames]
        gptrs = (C.c_void_p * 2)(*[g.ctypes.data for g in grads])
        cpu_lib.b2fcpu_costvol_backward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        assert cpu_lib.b2fcpu_costvol_backward(ptrs, 2, B, Cn, h, w, 9, fwd, C.c_void_p(wide[:, off:].ctypes.data),
                                               wide.strides[0] // 4, gptrs) == 0
        assert o.rel_err(grads[0], G[name + "_gradref"]) < 1e-5 and o.rel_err(grads[1], G[name + "_gradframe"]) < 1e-5
    img = np.ascontiguousarray(G["frame_fut"].transpose(0, 2, 3, 1))
    grid, go = np.ascontiguousarray(G["grid_fut"]), np.ascontiguousarray(G["warp_gradout3"])
    B, H, W, Cn = img.shape
    out = np.empty_like(img)
    assert cpu_lib.b2fcpu_warp_forward(_fp(img), _fp(grid), _fp(out), B, H, W, Cn, H, W) == 0
    assert o.rel_err(out, G["warp_img_fut"]) < 1e-5
    gi, gg = np.zeros_like(img), np.empty_like(grid)
    assert cpu_lib.b2fcpu_warp_backward(_fp(img), _fp(grid), _fp(go), _fp(gi), _fp(gg), B, H, W, Cn, H, W) == 0
    assert o.rel_err(gi, G["warp_img_fut_gradimg"]) < 1e-5 and o.rel_err(gg, G["warp_img_fut_gradgrid"]) < 1e-5


# ---------------------------------------------------------------------------------------
# ref_sampler_golden.npz: outputs of the REFERENCE'S OWN CUDA sampler (oracle/_ref/libstn_ref.so =
# extras/stnbhwd/BilinearSamplerBHWD.cu compiled unmodified, run on a B200 by
# tests/golden/make_ref_sampler_golden.py).  These pin the sampler oracles to the reference itself.
# ---------------------------------------------------------------------------------------

REF_PATH = os.path.join(ROOT, "tests", "golden", "ref_sampler_golden.npz")


@pytest.fixture(scope="module")
def RG():
    with np.load(REF_PATH) as z:
        return {k: z[k] for k in z.files}


def _ref_cases(RG):
    return sorted({k.split("__")[0] for k in RG})


def test_reference_sampler_fixture_covers_the_delicate_cases(RG):
    names = _ref_cases(RG)
    assert names == ["far_out", "feat32", "feat96", "img3_sigma4", "img3_sub", "one_pixel", "smallgrid_int"]
    for n in names:
        img, grid, out = RG[n + "__img"], RG[n + "__grid"], RG[n + "__out"]
        assert out.shape == grid.shape[:3] + img.shape[3:] and out.dtype == np.float32
        assert np.isfinite(out).all() and np.isfinite(RG[n + "__gradimg"]).all() and np.isfinite(RG[n + "__gradgrid"]).all()
    g = RG["far_out__grid"]
    x = np.arange(g.shape[2], dtype=np.float32)[None, None, :] + g[..., 0]
    assert (x < 0).any() and (x > g.shape[2] - 1).any()          # the clamp is exercised on both sides
    assert (RG["smallgrid_int__grid"] == np.round(RG["smallgrid_int__grid"])).mean() > 0.9   # exact-integer coordinates


def test_numpy_sampler_oracle_matches_the_reference_kernel(RG):
    for n in _ref_cases(RG):
        img, grid, go = RG[n + "__img"], RG[n + "__grid"], RG[n + "__gradout"]
        assert o.rel_err(o.warp_forward(img, grid), RG[n + "__out"]) < 1e-6, n
        gi, gg = o.warp_backward(img, grid, go)
        assert o.rel_err(gg, RG[n + "__gradgrid"]) < 2e-6, n
        assert o.rel_err(gi, RG[n + "__gradimg"]) < 1e-5, n            # float atomics in the reference
        _, gg_only = o.warp_backward(img, grid, go, only_grid=True)
        assert np.array_equal(gg_only, gg), n
    n = "img3_sigma4"                                                  # the independent scalar restatement too
    assert o.rel_err(o.warp_forward_loops(RG[n + "__img"], RG[n + "__grid"]), RG[n + "__out"]) < 1e-6


def test_c_sampler_oracle_matches_the_reference_kernel(RG, cpu_lib):
    for n in _ref_cases(RG):
        img, grid, go = (np.ascontiguousarray(RG[n + k]) for k in ("__img", "__grid", "__gradout"))
        B, H, W, Cn = img.shape
        _, Hg, Wg, _ = grid.shape
        out = np.empty((B, Hg, Wg, Cn), np.float32)
        assert cpu_lib.b2fcpu_warp_forward(_fp(img), _fp(grid), _fp(out), B, H, W, Cn, Hg, Wg) == 0
        assert o.rel_err(out, RG[n + "__out"]) < 1e-6, n
        gi, gg = np.zeros_like(img), np.empty_like(grid)
        assert cpu_lib.b2fcpu_warp_backward(_fp(img), _fp(grid), _fp(go), _fp(gi), _fp(gg), B, H, W, Cn, Hg, Wg) == 0
        assert o.rel_err(gg, RG[n + "__gradgrid"]) < 2e-5 and o.rel_err(gi, RG[n + "__gradimg"]) < 1e-5, n
