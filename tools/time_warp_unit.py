#!/usr/bin/env python
"""Row N2: the reference's warpingUnit chain (Transpose x2 -> MulConstant -> BilinearSamplerBHWD -> Transpose; the
layout / scaling passes done by torch copies standing in for THC's) against the fused BDHW entry points, at the
BASELINE config-2 shapes, for a smooth flow field (what the decoder produces) and an i.i.d. one."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from back2future_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
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
g = torch.Generator(device="cuda").manual_seed(2)
B = 8
for (Cn, H, W, scale) in ((32, 112, 256, 5.0), (64, 56, 128, 2.5), (128, 14, 32, 0.625), (3, 448, 1024, 20.0), (3, 224, 512, 10.0)):
    img = torch.randn(B, Cn, H, W, device=dev, generator=g)
    go = torch.randn(B, Cn, H, W, device=dev, generator=g)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    k = W / 1024.0
    smooth = torch.stack([6 * k * torch.sin(xs / (90.0 * k) + ys / (70.0 * k)) + 3 * k, 5 * k * torch.cos(xs / (60.0 * k) - ys / (110.0 * k))], 0)
    flows = {"smooth": ((smooth[None] + 0.05 * torch.randn(B, 2, H, W, device=dev, generator=g)) / scale).contiguous(),
             "iid4": (torch.randn(B, 2, H, W, device=dev, generator=g) * 4 / scale).contiguous()}
    for name, flow in flows.items():
        out, gi, gf = torch.empty_like(img), torch.zeros_like(img), torch.empty_like(flow)
        img_t, out_t, gi_t = (torch.empty(B, H, W, Cn, device=dev) for _ in range(3))
        grid_t, gg_t = torch.empty(B, H, W, 2, device=dev), torch.empty(B, H, W, 2, device=dev)
        def chain_fwd():
            img_t.copy_(img.permute(0, 2, 3, 1)); grid_t.copy_((flow * scale).permute(0, 2, 3, 1))
            _lib.check(lib.b2f_warp_bhwd_forward(P(img_t), P(grid_t), P(out_t), B, H, W, Cn, H, W, None))
            out.copy_(out_t.permute(0, 3, 1, 2))
        def chain_bwd():
            go_t = go.permute(0, 2, 3, 1).contiguous(); gi_t.zero_()
            _lib.check(lib.b2f_warp_bhwd_backward(P(img_t), P(grid_t), P(go_t), P(gi_t), P(gg_t), B, H, W, Cn, H, W, None))
            gi.copy_(gi_t.permute(0, 3, 1, 2)); gf.copy_(gg_t.permute(0, 3, 1, 2) * scale)
        def fused_fwd():
            _lib.check(lib.b2f_warp_bdhw_forward(P(img), P(flow), scale, P(out), B, Cn, H, W, None))
        def fused_bwd():
            gi.zero_()
            _lib.check(lib.b2f_warp_bdhw_backward(P(img), P(flow), scale, P(go), P(gi), P(gf), B, Cn, H, W, None))
        cf, cb, ff, fb = timeit(chain_fwd), timeit(chain_bwd), timeit(fused_fwd), timeit(fused_bwd)
        print("8x%-3dx%dx%-4d %-6s chain fwd %6.1f bwd %6.1f us | fused fwd %6.1f bwd %6.1f us | x%.2f / x%.2f"
              % (Cn, H, W, name, cf, cb, ff, fb, cf / ff, cb / fb))
