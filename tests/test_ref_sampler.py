"""GPU parity of the sampler against the REFERENCE'S OWN CUDA code (rows a4/a5 of SURVEY.md section 8).

oracle/_ref/libstn_ref.so = the reference's unmodified extras/stnbhwd/BilinearSamplerBHWD.cu (Lua-C glue
and kernels) compiled against stand-in Torch7 headers (oracle/ref_shim/).  Three-way check on the same
seeded inputs: reference kernel vs numpy oracle (pins the oracle), reference kernel vs the CUDA product
through the C ABI (the parity claim itself), and the committed golden file (which is this library's
output) vs what the library produces on this box.

Tolerances: output and flow gradient 1e-6 relative -- same fp32 geometry, products summed in a different
order / with different FMA contraction only; image gradient 1e-5 (float atomics: the reference's own sum
order varies run to run).  The north_star bar is 1e-4.
"""
import os

import numpy as np
import pytest

from oracle import b2f_oracle as o
from oracle import ref_sampler

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref(cuda_lib):
    if not ref_sampler.available():
        pytest.skip("oracle/_ref/libstn_ref.so not built (needs /root/reference at build time)")
    ref_sampler.load()
    return ref_sampler


def _product(cuda_lib, img, grid, go, only_grid=False):
    import ctypes as C
    import torch
    from back2future_b200 import _lib
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)   # noqa: E731
    p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None                   # noqa: E731
    ti, tg, tgo = t(img), t(grid), t(go)
    B, H, W, Cn = ti.shape
    _, Hg, Wg, _ = tg.shape
    out = torch.full((B, Hg, Wg, Cn), float("nan"), device=dev)
    gi = None if only_grid else torch.zeros_like(ti)
    gg = torch.full_like(tg, float("nan"))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(cuda_lib.b2f_warp_bhwd_forward(p(ti), p(tg), p(out), B, H, W, Cn, Hg, Wg, st))
    _lib.check(cuda_lib.b2f_warp_bhwd_backward(p(ti), p(tg), p(tgo), p(gi), p(gg), B, H, W, Cn, Hg, Wg, st))
    torch.cuda.synchronize()
    return out.cpu().numpy(), (None if gi is None else gi.cpu().numpy()), gg.cpu().numpy()


@pytest.mark.parametrize("B,H,W,Cn,Hg,Wg,sigma", [
    (2, 14, 32, 128, 14, 32, 4.0), (2, 28, 64, 96, 28, 64, 4.0), (1, 17, 23, 32, 17, 23, 0.5),
    (2, 9, 11, 3, 9, 11, 4.0), (1, 33, 65, 3, 33, 65, 0.5), (1, 6, 7, 5, 6, 7, 2.0), (1, 5, 6, 1, 5, 6, 2.0),
    (1, 4, 5, 4, 4, 5, 30.0), (3, 1, 1, 8, 1, 1, 1.0), (1, 16, 16, 192, 16, 16, 3.0), (2, 9, 12, 8, 5, 7, 3.0),
    (1, 56, 128, 64, 56, 128, 4.0), (2, 112, 256, 3, 112, 256, 4.0),
    (2, 12, 16, 3, 12, 16, 4.0), (1, 24, 64, 3, 24, 64, 30.0), (2, 9, 12, 3, 5, 8, 3.0), (1, 1, 4, 3, 1, 4, 1.0),
])
def test_reference_kernel_vs_oracle_vs_product(ref, cuda_lib, B, H, W, Cn, Hg, Wg, sigma):
    r = np.random.default_rng(11)
    img = r.standard_normal((B, H, W, Cn)).astype(np.float32)
    grid = (r.standard_normal((B, Hg, Wg, 2)) * sigma).astype(np.float32)
    go = r.standard_normal((B, Hg, Wg, Cn)).astype(np.float32)
    rout, rgi, rgg = ref.run(img, grid, go)
    # the oracle restates the reference
    ogi, ogg = o.warp_backward(img, grid, go)
    assert o.rel_err(o.warp_forward(img, grid), rout) < 1e-6
    assert o.rel_err(ogg, rgg) < 2e-6 and o.rel_err(ogi, rgi) < 1e-5
    # the product reproduces the reference
    pout, pgi, pgg = _product(cuda_lib, img, grid, go)
    assert o.rel_err(pout, rout) < 1e-6
    assert o.rel_err(pgg, rgg) < 2e-6 and o.rel_err(pgi, rgi) < 1e-5
    # the reference's never-called OnlyGrid entry (.cu:368-419) == our gradImg == NULL mode
    _, none, rgg_only = ref.run(img, grid, go, only_grid=True)
    _, pnone, pgg_only = _product(cuda_lib, img, grid, go, only_grid=True)
    assert none is None and pnone is None
    assert np.array_equal(rgg_only, rgg) and np.array_equal(pgg_only, pgg)


def test_reference_kernel_integer_and_border_coordinates(ref, cuda_lib):
    """floor / clamp decisions (which cell, which taps are 'in') must agree exactly: with exact-integer,
    exactly-on-the-border and just-off-zero coordinates any disagreement shows up as an O(1) error."""
    r = np.random.default_rng(12)
    img = r.standard_normal((2, 9, 12, 8)).astype(np.float32)
    grid = np.round(r.standard_normal((2, 9, 12, 2)) * 3).astype(np.float32)
    grid[0, 0, 0] = (11.0, 8.0)
    grid[0, 1, 1] = (-1e-8, 1e-8)
    grid[1, 8, 11] = (0.0, 0.0)
    grid[1, 8, 10] = (1.0, 0.0)
    grid[1, 3, 3] = (7.9999995, 4.9999995)
    go = r.standard_normal((2, 9, 12, 8)).astype(np.float32)
    rout, rgi, rgg = ref.run(img, grid, go)
    pout, pgi, pgg = _product(cuda_lib, img, grid, go)
    assert o.rel_err(pout, rout) < 1e-6 and o.rel_err(pgg, rgg) < 2e-6 and o.rel_err(pgi, rgi) < 1e-5
    assert o.rel_err(o.warp_forward(img, grid), rout) < 1e-6


def test_golden_file_is_what_the_reference_produces_here(ref):
    """tests/golden/ref_sampler_golden.npz was written by make_ref_sampler_golden.py on a B200; the same
    library must reproduce it on this box (bit-exact except the atomically accumulated gradImg)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_sampler_golden.npz")
    with np.load(path) as z:
        G = {k: z[k] for k in z.files}
    names = sorted({k.split("__")[0] for k in G})
    assert len(names) >= 7
    for n in names:
        out, gi, gg = ref.run(G[n + "__img"], G[n + "__grid"], G[n + "__gradout"])
        assert np.array_equal(out, G[n + "__out"]) and np.array_equal(gg, G[n + "__gradgrid"]), n
        assert o.rel_err(gi, G[n + "__gradimg"]) < 1e-5, n


def test_product_reproduces_reference_golden(cuda_lib):
    """Runs without the reference library: the committed reference outputs vs the CUDA product."""
    path = os.path.join(ROOT, "tests", "golden", "ref_sampler_golden.npz")
    with np.load(path) as z:
        G = {k: z[k] for k in z.files}
    for n in sorted({k.split("__")[0] for k in G}):
        out, gi, gg = _product(cuda_lib, G[n + "__img"], G[n + "__grid"], G[n + "__gradout"])
        assert o.rel_err(out, G[n + "__out"]) < 1e-6, n
        assert o.rel_err(gg, G[n + "__gradgrid"]) < 2e-6 and o.rel_err(gi, G[n + "__gradimg"]) < 1e-5, n
