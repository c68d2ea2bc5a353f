"""One tensor-core convolution of the level-3 decoder (128 -> 128 channels at 112 x 256, the shape that dominates the
whole-network forward) in a loop, for ncu:  ncu --set full -k regex:conv3x3_tc -s 3 -c 2 python tools/prof_tc.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from back2future_b200 import _lib

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Cin = Cout = 128
H, W = 112, 256
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
xh = torch.randn(B, H, W, Cin, device="cuda")
xl = torch.randn(B, H, W, Cin, device="cuda") * 1e-4
n = int(lib.b2f_conv3x3_tc_packed_floats(Cin, Cout))
wh, wl = torch.randn(n, device="cuda") * 0.03, torch.randn(n, device="cuda") * 1e-5
oh, ol = torch.empty(B, H, W, Cout, device="cuda"), torch.empty(B, H, W, Cout, device="cuda")
bias = torch.zeros(Cout, device="cuda")
for _ in range(6):
    _lib.check(lib.b2f_conv3x3_tc_forward(p(xh), p(xl), p(wh), p(wl), p(bias), p(oh), p(ol), None, 0, B, Cin, H, W, Cout, 0.2, st))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    lib.b2f_conv3x3_tc_forward(p(xh), p(xl), p(wh), p(wl), p(bias), p(oh), p(ol), None, 0, B, Cin, H, W, Cout, 0.2, st)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
fl = 2.0 * B * H * W * Cin * Cout * 9
print("conv3x3_tc %dx%d -> %d, B=%d, %dx%d: %.3f ms, %.1f TFLOP/s fp32-equivalent, %.1f TFLOP/s of TF32 MMA (3 passes)"
      % (Cin, Cin, Cout, B, H, W, ms, fl / ms / 1e9, 3 * fl / ms / 1e9))
