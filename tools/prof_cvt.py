"""Level-3 cost-volume forward on both paths for ncu: the tensor-core kernel (b2f_debug_costvol_path 16) and the FFMA2
kernel (17).  ncu --set full -k regex:costvol_fwd -s 4 -c 4 python tools/prof_cvt.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from back2future_b200 import _lib
lib = _lib.load()
B, Cn, h, w = 8, 32, 112, 256
P = lambda t: C.c_void_p(t.data_ptr())
ref, frm = torch.randn(B, Cn, h, w, device="cuda"), torch.randn(B, Cn, h, w, device="cuda")
out = torch.empty(B, 162, h, w, device="cuda")
fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()])
for rep in range(4):
    for m in (16, 17):
        lib.b2f_debug_costvol_path(m)
        _lib.check(lib.b2f_costvol_forward(fp, 2, B, Cn, h, w, 9, 1, P(out), out.stride(0), None))
lib.b2f_debug_costvol_path(0)
torch.cuda.synchronize()
