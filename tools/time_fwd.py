#!/usr/bin/env python
"""Time the cost-volume forward/backward at levels 3 and 4 for the library in B2F_LIB_PATH (kernel experiments)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from back2future_b200 import _lib
lib = _lib.load()
if os.environ.get('B2F_PATH'):
    lib.b2f_debug_costvol_path(int(os.environ['B2F_PATH']))
dev = torch.device("cuda:0")
B = 8
P = lambda t: C.c_void_p(t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=12):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
what = sys.argv[1] if len(sys.argv) > 1 else "fb"
row = "%-10s" % os.path.basename(os.environ.get("B2F_LIB_PATH", "default"))[:10]
for l, Cn in ((3, 32), (4, 64), (5, 96)):
    h, w = 448 >> (l - 1), 1024 >> (l - 1)
    ref, frm = torch.randn(B, Cn, h, w, device=dev), torch.randn(B, Cn, h, w, device=dev)
    joined, gj = torch.empty(B, 162, h, w, device=dev), torch.randn(B, 162, h, w, device=dev)
    gr, gf = torch.empty_like(ref), torch.empty_like(frm)
    fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()]); gp = _lib.ptr_array([gr.data_ptr(), gf.data_ptr()])
    row += " | L%d" % l
    if "f" in what:
        for sg in (1, 0):
            row += " f%d %6.1f" % (sg, timeit(lambda: _lib.check(lib.b2f_costvol_forward(fp, 2, B, Cn, h, w, 9, sg, P(joined), joined.stride(0), None))))
    if "b" in what:
        row += " b %6.1f" % timeit(lambda: _lib.check(lib.b2f_costvol_backward(fp, 2, B, Cn, h, w, 9, 1, P(gj), gj.stride(0), gp, None)))
print(row)
