"""ctypes binding of oracle/c/b2f_check64.c -- TEST INFRASTRUCTURE, NOT PRODUCT.

The float64 closed-form checker of the cost volume and the sampler, fast enough (OpenMP) to compare EVERY op of the
benchmarked step at its full size.  `tests/test_oracle.py` pins it to the numpy oracle (oracle/b2f_oracle.py) on
small ragged cases; only tests/, smoke() and bench.py's checker legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "c", "libb2f_check64.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            import subprocess
            env = dict(os.environ)
            env.pop("CC", None)
            subprocess.run(["make", "-C", os.path.dirname(_HERE), "oracle/c/libb2f_check64.so"], check=True, env=env,
                           capture_output=True)
        lib = C.CDLL(LIB_PATH)
        lib.b2fchk_rel_err.restype = C.c_double
        lib.b2fchk_rel_err.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]
        lib.b2fchk_costvol_backward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _vp(a):
    return C.c_void_p(a.ctypes.data)


def costvol_forward(frames, win=9, fwd=True):
    lib = load()
    fr = [_f32(f) for f in frames]
    B, Cn, H, W = fr[0].shape
    out = np.empty((B, win * win, H, W), np.float64)
    ptrs = (C.c_void_p * len(fr))(*[f.ctypes.data for f in fr])
    assert lib.b2fchk_costvol_forward(ptrs, len(fr), B, Cn, H, W, win, int(bool(fwd)), _vp(out)) == 0
    return out


def costvol_backward(frames, grad_out, win=9, fwd=True):
    """grad_out (B, win*win, H, W), possibly a batch-strided view of a wider float32 array."""
    lib = load()
    fr = [_f32(f) for f in frames]
    B, Cn, H, W = fr[0].shape
    go = np.asarray(grad_out)
    if go.dtype != np.float32 or go.strides[1:] != (H * W * 4, W * 4, 4):
        go = _f32(go)
    grads = [np.empty((B, Cn, H, W), np.float64) for _ in fr]
    ptrs = (C.c_void_p * len(fr))(*[f.ctypes.data for f in fr])
    gptrs = (C.c_void_p * len(fr))(*[g.ctypes.data for g in grads])
    assert lib.b2fchk_costvol_backward(ptrs, len(fr), B, Cn, H, W, win, int(bool(fwd)), _vp(go),
                                       go.strides[0] // 4 if B > 1 else 0, gptrs) == 0
    return grads


def warp_forward(img, grid):
    lib = load()
    img, grid = _f32(img), _f32(grid)
    B, H, W, Cn = img.shape
    _, Hg, Wg, _ = grid.shape
    out = np.empty((B, Hg, Wg, Cn), np.float64)
    assert lib.b2fchk_warp_forward(_vp(img), _vp(grid), _vp(out), B, H, W, Cn, Hg, Wg) == 0
    return out


def warp_backward(img, grid, grad_out, only_grid=False):
    lib = load()
    img, grid, go = _f32(img), _f32(grid), _f32(grad_out)
    B, H, W, Cn = img.shape
    _, Hg, Wg, _ = grid.shape
    gimg = None if only_grid else np.empty((B, H, W, Cn), np.float64)
    ggrid = np.empty((B, Hg, Wg, 2), np.float64)
    assert lib.b2fchk_warp_backward(_vp(img), _vp(grid), _vp(go), _vp(gimg) if gimg is not None else None,
                                    _vp(ggrid), B, H, W, Cn, Hg, Wg) == 0
    return gimg, ggrid


def rel_err(kernel_f32, ref_f64):
    """max |a-b| / max(|b|, rms(b)); `kernel_f32` may be a view whose rows (last three dims flattened per leading
    index) are dense -- e.g. one half of the 162-channel joined buffer."""
    lib = load()
    a = np.asarray(kernel_f32)
    b = np.ascontiguousarray(ref_f64, dtype=np.float64)
    assert a.dtype == np.float32 and a.shape == b.shape
    if a.ndim == 4 and not a.flags.c_contiguous:
        inner = a.shape[1] * a.shape[2] * a.shape[3]
        assert a.strides[1:] == (a.shape[2] * a.shape[3] * 4, a.shape[3] * 4, 4)
        rows, row, rstride = a.shape[0], inner, a.strides[0] // 4
    else:
        a = np.ascontiguousarray(a)
        rows, row, rstride = 1, a.size, a.size
    worst = C.c_int64(-1)
    return float(lib.b2fchk_rel_err(_vp(a), rows, row, rstride, _vp(b), C.byref(worst)))
