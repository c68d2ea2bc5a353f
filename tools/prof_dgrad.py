"""The tensor-core input gradients in a loop, for ncu: the level-3 128 -> 128 decoder layer with the channel-minor mask
(conv3x3_tc_kernel<128, 2, false>) and the stride-2 form of the 64 -> 32 pyramid layer (conv3x3_tc_kernel<32, 0, true>):
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -c 12 python tools/prof_dgrad.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from back2future_b200 import _lib

lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


B, H, W, Cn = 8, 80, 160, 128
gh, gl = torch.randn(B, H, W, Cn, device="cuda"), torch.randn(B, H, W, Cn, device="cuda") * 1e-4
ah = torch.randn(B, H, W, Cn, device="cuda")
nt = 9 * Cn * Cn
th, tl = torch.randn(nt, device="cuda") * 0.03, torch.randn(nt, device="cuda") * 1e-5
oh, ol = torch.empty(B, H, W, Cn, device="cuda"), torch.empty(B, H, W, Cn, device="cuda")
ms = timed(lambda: _lib.check(lib.b2f_conv3x3_tc_backward_data(p(gh), p(gl), p(th), p(tl), None, 0, p(ah), p(oh), p(ol), None, 0,
                                                               B, Cn, H, W, Cn, 0.2, 0, st)), 3)
fl = 2.0 * B * H * W * Cn * Cn * 9
print("input gradient 128 -> 128 at %d x %d x %d: %.3f ms, %.1f TFLOP/s of TF32 MMA" % (B, H, W, ms, 3 * fl / ms / 1e9))

B2, Ho, Wo, Co, Ci = 24, 40, 80, 64, 32
g2h, g2l = torch.randn(B2, Ho, Wo, Co, device="cuda"), torch.randn(B2, Ho, Wo, Co, device="cuda") * 1e-4
n2 = 9 * Ci * Co
t2h, t2l = torch.randn(n2, device="cuda") * 0.03, torch.randn(n2, device="cuda") * 1e-5
gin = torch.zeros(B2, Ci, 2 * Ho, 2 * Wo, device="cuda")
ms = timed(lambda: _lib.check(lib.b2f_conv3x3_tc_backward_data_s2(p(g2h), p(g2l), p(t2h), p(t2l), p(gin), 0, B2, Co, Ho, Wo, Ci,
                                                                  2 * Ho, 2 * Wo, 1, st)), 3)
print("stride-2 input gradient 64 -> 32 at %d x %d x %d (output %d x %d): %.3f ms" % (B2, Ho, Wo, 2 * Ho, 2 * Wo, ms))
