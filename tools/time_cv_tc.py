#!/usr/bin/env python
"""Tensor-core cost-volume forward (b2f_debug_costvol_path 16) against the FFMA2 kernel (path 17 = forbid): parity
against the float64 checker on small / ragged shapes, agreement and timing at the pyramid levels."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from back2future_b200 import _lib
from oracle import check64 as c64
from oracle import b2f_oracle as o
lib = _lib.load()
dev = torch.device("cuda:0")
P = lambda t: C.c_void_p(t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def run(mode, ref, frm, fwd, out=None):
    B, Cn, h, w = ref.shape
    if out is None:
        out = torch.full((B, 81, h, w), float("nan"), device=dev)
    fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()])
    lib.b2f_debug_costvol_path(mode)
    try:
        _lib.check(lib.b2f_costvol_forward(fp, 2, B, Cn, h, w, 9, int(fwd), P(out), out.stride(0), None))
    finally:
        lib.b2f_debug_costvol_path(0)
    torch.cuda.synchronize()
    return out


rng = np.random.default_rng(2)
bad = 0
for (B, Cn, h, w) in ((1, 32, 8, 16), (2, 32, 16, 32), (1, 64, 24, 48), (2, 40, 13, 20), (1, 96, 9, 72), (3, 3, 5, 8), (1, 192, 7, 16)):
    ref = rng.standard_normal((B, Cn, h, w)).astype(np.float32)
    frm = rng.standard_normal((B, Cn, h, w)).astype(np.float32)
    for fwd in (True, False):
        got = run(16, torch.from_numpy(ref).to(dev), torch.from_numpy(frm).to(dev), fwd).cpu().numpy()
        want = c64.costvol_forward([ref, frm], 9, fwd)
        e = o.rel_err(got, want)
        print("parity B=%d C=%3d %3dx%-3d fwd=%d  rel_err %.2e %s" % (B, Cn, h, w, fwd, e, "ok" if e < 1e-4 else "FAIL"))
        bad += not (e < 1e-4)

B = 8
for l, Cn in ((3, 32), (4, 64), (5, 96), (6, 128), (7, 192)):
    h, w = 448 >> (l - 1), 1024 >> (l - 1)
    ref, frm = torch.randn(B, Cn, h, w, device=dev), torch.randn(B, Cn, h, w, device=dev)
    joined = torch.empty(B, 162, h, w, device=dev)
    fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()])
    row = "L%d C=%3d %3dx%-4d" % (l, Cn, h, w)
    outs = {}
    for m in (17, 16):
        outs[m] = run(m, ref, frm, True)
        lib.b2f_debug_costvol_path(m)
        t = timeit(lambda: _lib.check(lib.b2f_costvol_forward(fp, 2, B, Cn, h, w, 9, 1, P(joined), joined.stride(0), None)))
        lib.b2f_debug_costvol_path(0)
        alg = 4 * B * h * w * (2 * Cn + 81)
        row += " | %s %7.1f us %6.0f GB/s" % ("ffma2" if m == 17 else "tc   ", t, alg / t / 1e3)
    d = (outs[16] - outs[17]).abs().max().item() / outs[17].abs().max().item()
    row += " | max diff / max %.1e" % d
    bad += not (d < 1e-4)
    print(row)
print("FAILED" if bad else "all ok")
sys.exit(1 if bad else 0)
