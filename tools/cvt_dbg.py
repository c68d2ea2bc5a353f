import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
import torch
from back2future_b200 import _lib
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
B=8
for l, Cn in ((3, 32), (4, 64)):
    h, w = 448 >> (l - 1), 1024 >> (l - 1)
    ref, frm = torch.randn(B, Cn, h, w, device=dev), torch.randn(B, Cn, h, w, device=dev)
    out = torch.empty(B, 81, h, w, device=dev)
    fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()])
    row = "L%d" % l
    for m in (16, 18, 19):
        lib.b2f_debug_costvol_path(m)
        t = timeit(lambda: _lib.check(lib.b2f_costvol_forward(fp, 2, B, Cn, h, w, 9, 1, P(out), out.stride(0), None)))
        lib.b2f_debug_costvol_path(0)
        row += " | mode %d %6.1f us" % (m, t)
    print(row)
