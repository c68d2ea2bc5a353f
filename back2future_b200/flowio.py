"""Middlebury `.flo` files -- host-side mirror of flowExtensions.lua:254-287 (`loadFLO`, `writeFLO`), the wire
format of the flow fields `computeFlow` returns (README.md:49-71).  Thin wrappers over the C ABI
(`b2f_flo_write / b2f_flo_read_header / b2f_flo_read`); arrays are numpy float32 in the reference's planar
(2, h, w) layout."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def writeFLO(filename, F):
    """flowExtensions.lua:274-286: F is (2, h, w) -- channel 0 = u, 1 = v."""
    F = np.ascontiguousarray(F, dtype=np.float32)
    if F.ndim != 3 or F.shape[0] != 2:
        raise ValueError("writeFLO: expected a (2, h, w) array, got %r" % (F.shape,))
    lib = _lib.load()
    _lib.check(lib.b2f_flo_write(str(filename).encode(), F.ctypes.data_as(C.c_void_p), F.shape[1], F.shape[2]))


def loadFLO(filename):
    """flowExtensions.lua:254-271: returns the (2, h, w) float32 flow; a wrong tag raises like the reference's
    'unable to read ... perhaps bigendian error'."""
    lib = _lib.load()
    w, h = C.c_int(0), C.c_int(0)
    _lib.check(lib.b2f_flo_read_header(str(filename).encode(), C.byref(w), C.byref(h)))
    out = np.empty((2, h.value, w.value), np.float32)
    _lib.check(lib.b2f_flo_read(str(filename).encode(), out.ctypes.data_as(C.c_void_p), h.value, w.value))
    return out
