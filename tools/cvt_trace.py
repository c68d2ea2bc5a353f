"""clock64 timeline of the tensor-core cost-volume forward (costvol_tc.cu; buffer registered with b2f_debug_tc_trace):
per tile of the first 8 CTAs, when the converters finished each operand group, when the MMA issuer started each chunk,
when the accumulators were complete and when the epilogue was done.  usage: python tools/cvt_trace.py [level=3]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from back2future_b200 import _lib
lib = _lib.load()
l = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Cn = {3: 32, 4: 64, 5: 96, 6: 128, 7: 192}[l]
B, h, w = 8, 448 >> (l - 1), 1024 >> (l - 1)
P = lambda t: C.c_void_p(t.data_ptr())
ref, frm = torch.randn(B, Cn, h, w, device="cuda"), torch.randn(B, Cn, h, w, device="cuda")
out = torch.empty(B, 81, h, w, device="cuda")
fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()])
lib.b2f_debug_costvol_path(16)
run = lambda: _lib.check(lib.b2f_costvol_forward(fp, 2, B, Cn, h, w, 9, 1, P(out), out.stride(0), None))
for _ in range(3):
    run()
buf = torch.zeros(8 * 8 * 16, device="cuda", dtype=torch.int64)
lib.b2f_debug_tc_trace(P(buf))
run()
torch.cuda.synchronize()
lib.b2f_debug_tc_trace(None)
lib.b2f_debug_costvol_path(0)
t = buf.cpu().numpy().reshape(8, 8, 16).astype(np.int64)
names = ["cvt A", "cvt B0", "cvt B1", "cvt B2", "mma c0", "mma c1", "mma c2", "mma issued", "acc full", "epi done", "raw A", "raw B0", "raw B1", "raw B2"]
for cta in range(2):
    t0 = t[cta, 0][t[cta, 0] > 0].min()
    print("CTA %d (cycles since its first stamp)" % cta)
    print("  tile " + " ".join("%10s" % n for n in names))
    for tl in range(8):
        print("  %4d " % tl + " ".join("%10d" % (t[cta, tl, k] - t0 if t[cta, tl, k] else -1) for k in range(14)))
