"""One weight-gradient call of the level-3 decoder (128 -> 128 channels at 80 x 160, B = 8: the training shape) in a
loop, for ncu:  ncu --set full -k regex:wgrad -s 2 -c 1 python tools/prof_wgrad.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from back2future_b200 import _lib
lib = _lib.load()
B, Cin, Cout, H, W = 8, 128, 128, 80, 160
p = lambda t: C.c_void_p(t.data_ptr())
x, g = torch.randn(B, Cin, H, W, device="cuda"), torch.randn(B, Cout, H, W, device="cuda")
n = int(lib.b2f_conv3x3_packed_floats(Cin, Cout))
gw, gb = torch.zeros(n, device="cuda"), torch.zeros(Cout, device="cuda")
run = lambda: _lib.check(lib.b2f_conv3x3_backward_weights(p(x), 0, p(g), 0, p(gw), p(gb), B, Cin, H, W, Cout, 1, None))
for _ in range(3):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("wgrad %d -> %d, B=%d, %dx%d: %.3f ms, %.1f TFLOP/s" % (Cin, Cout, B, H, W, ms, 2.0 * B * H * W * Cin * Cout * 9 / ms / 1e9))

# the same layer on the tensor cores
cxp = (Cin + 31) // 32 * 32
xh, xl = torch.empty(B, H, W, cxp, device="cuda"), torch.empty(B, H, W, cxp, device="cuda")
gh, gl = torch.empty(B, H, W, Cout, device="cuda"), torch.empty(B, H, W, Cout, device="cuda")
_lib.check(lib.b2f_nhwc_split_from_bdhw(p(x), 0, p(xh), p(xl), B, Cin, H, W, None))
_lib.check(lib.b2f_nhwc_split_from_bdhw(p(g), 0, p(gh), p(gl), B, Cout, H, W, None))
run_tc = lambda: _lib.check(lib.b2f_conv3x3_tc_backward_weights(p(xh), p(xl), Cin, p(gh), p(gl), p(g), 0, p(gw), p(gb), B, Cin, H, W, Cout, None))
for _ in range(3):
    run_tc()
torch.cuda.synchronize()
a.record()
for _ in range(10):
    run_tc()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("wgrad_tc %d -> %d, B=%d, %dx%d: %.3f ms, %.1f TFLOP/s fp32-equivalent" % (Cin, Cout, B, H, W, ms, 2.0 * B * H * W * Cin * Cout * 9 / ms / 1e9))
