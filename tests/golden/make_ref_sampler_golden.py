#!/usr/bin/env python
"""Generates tests/golden/ref_sampler_golden.npz  --  run ON A GPU BOX (the reference's sampler is CUDA):

    gpurun -- 'python tests/golden/make_ref_sampler_golden.py gpurun_out/ref_sampler_golden.npz'

and copy the result to tests/golden/.  Unlike hotpath_golden.npz these expected values are produced by
the REFERENCE ITSELF: oracle/_ref/libstn_ref.so is the reference's unmodified
extras/stnbhwd/BilinearSamplerBHWD.cu (Lua-C glue + kernels) compiled against stand-in Torch7 headers
(oracle/ref_shim/, oracle/ref_sampler.py).  They pin rows a4/a5 of SURVEY.md section 8 (quirks Q1-Q3): the
numpy and C oracles are checked against them on the CPU (tests/test_golden.py), the CUDA product on the
GPU (tests/test_ref_sampler.py).

Cases cover what the reference's semantics make delicate: C = 3 (image warps) and C % 4 == 0 / C > 32
(feature warps, the kernel's `t += 32` channel loop), grid smaller than the image, sub-pixel, multi-pixel
and far-out-of-image flow (border clamp, the tap at index W / H that reads as zero, no clamp derivative),
exact-integer and exactly-on-the-border coordinates, width not a multiple of the 16-pixel block.
gradImg is accumulated with float atomics in the reference (sum order varies run to run): consumers compare
it at 1e-5, everything else at 1e-6.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_sampler  # noqa: E402

# name: (B, H, W, C, Hg, Wg, sigma, kind)
CASES = {
    "img3_sigma4": (2, 9, 11, 3, 9, 11, 4.0, "normal"),
    "img3_sub": (1, 20, 37, 3, 20, 37, 0.5, "normal"),
    "feat32": (1, 17, 23, 32, 17, 23, 2.0, "normal"),
    "feat96": (1, 7, 18, 96, 7, 18, 3.0, "normal"),
    "smallgrid_int": (2, 9, 12, 8, 5, 7, 3.0, "integer"),
    "far_out": (1, 4, 5, 4, 4, 5, 30.0, "normal"),
    "one_pixel": (3, 1, 1, 8, 1, 1, 1.0, "normal"),
}


def inputs(name):
    B, H, W, Cn, Hg, Wg, sigma, kind = CASES[name]
    r = np.random.default_rng(sum(map(ord, name)))
    img = r.standard_normal((B, H, W, Cn)).astype(np.float32)
    grid = (r.standard_normal((B, Hg, Wg, 2)) * sigma).astype(np.float32)
    if kind == "integer":
        grid = np.round(grid)
        grid[0, 0, 0] = (W - 1.0, H - 1.0)       # lands exactly on the last column / row
        grid[0, 1, 1] = (-1e-8, 1e-8)
        grid[0, 2, 2] = (0.5, -0.5)
    go = r.standard_normal((B, Hg, Wg, Cn)).astype(np.float32)
    return img, grid.astype(np.float32), go


def main(path):
    out = {}
    for name in CASES:
        img, grid, go = inputs(name)
        o, gi, gg = ref_sampler.run(img, grid, go)
        _, none, gg_only = ref_sampler.run(img, grid, go, only_grid=True)
        assert none is None and np.array_equal(gg, gg_only), name
        o2, gi2, gg2 = ref_sampler.run(img, grid, go)
        assert np.array_equal(o, o2) and np.array_equal(gg, gg2), name     # deterministic parts
        assert np.abs(gi - gi2).max() <= 1e-5 * max(1.0, np.abs(gi).max()), name
        for k, v in (("img", img), ("grid", grid), ("gradout", go), ("out", o), ("gradimg", gi), ("gradgrid", gg)):
            out["%s__%s" % (name, k)] = v
        print(name, "out", o.shape, "finite", np.isfinite(o).all() and np.isfinite(gi).all() and np.isfinite(gg).all())
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_sampler_golden.npz"))
