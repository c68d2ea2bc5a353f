#!/usr/bin/env python
"""Time cost-volume kernels at pyramid levels for each path mode (0 auto, 2/3/4 forced strip widths)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from back2future_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
B = 8
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
modes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 2, 3, 4]
for l, Cn in ((3, 32), (4, 64), (5, 96), (6, 128), (7, 192)):
    h, w = 448 >> (l - 1), 1024 >> (l - 1)
    ref, frm = torch.randn(B, Cn, h, w, device=dev), torch.randn(B, Cn, h, w, device=dev)
    joined, gj = torch.empty(B, 162, h, w, device=dev), torch.randn(B, 162, h, w, device=dev)
    gr, gf = torch.empty_like(ref), torch.empty_like(frm)
    fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()]); gp = _lib.ptr_array([gr.data_ptr(), gf.data_ptr()])
    row = "L%d C=%3d %3dx%-4d" % (l, Cn, h, w)
    for m in modes:
        lib.b2f_debug_costvol_path(m)
        f = timeit(lambda: _lib.check(lib.b2f_costvol_forward(fp, 2, B, Cn, h, w, 9, 1, P(joined), joined.stride(0), None)))
        bw = timeit(lambda: _lib.check(lib.b2f_costvol_backward(fp, 2, B, Cn, h, w, 9, 1, P(gj), gj.stride(0), gp, None)))
        row += " | m%d fwd %7.1f bwd %7.1f us" % (m, f, bw)
    lib.b2f_debug_costvol_path(0)
    print(row)
