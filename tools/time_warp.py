#!/usr/bin/env python
"""Time the C = 3 image warps (and one feature warp) for the library in B2F_LIB_PATH: full backward vs
flow-gradient-only, i.i.d. flow of sigma 4 / 0.5 px vs a smooth flow field (what the network produces)."""
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
print(os.path.basename(os.environ.get("B2F_LIB_PATH", "default")))
import os as _os
SHAPES = ((8, 112, 256, 32), (8, 56, 128, 64), (8, 28, 64, 96), (8, 14, 32, 128)) if _os.environ.get('TW_FEAT') else ((8, 448, 1024, 3), (8, 224, 512, 3), (8, 112, 256, 32))
for (B, H, W, Cn) in SHAPES:
    img = torch.randn(B, H, W, Cn, device=dev, generator=g)
    go = torch.randn(B, H, W, Cn, device=dev, generator=g)
    out, gi = torch.empty_like(img), torch.zeros_like(img)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    smooth = torch.stack([6 * torch.sin(xs / 90.0 + ys / 70.0) + 3, 5 * torch.cos(xs / 60.0 - ys / 110.0)], -1)
    flows = {"iid4": torch.randn(B, H, W, 2, device=dev, generator=g) * 4,
             "iid.5": torch.randn(B, H, W, 2, device=dev, generator=g) * 0.5,
             "smooth": (smooth[None] + 0.3 * torch.randn(B, H, W, 2, device=dev, generator=g)).contiguous()}
    for name, grid in flows.items():
        gg = torch.empty_like(grid)
        f = timeit(lambda: _lib.check(lib.b2f_warp_bhwd_forward(P(img), P(grid), P(out), B, H, W, Cn, H, W, None)))
        bw = timeit(lambda: _lib.check(lib.b2f_warp_bhwd_backward(P(img), P(grid), P(go), P(gi), P(gg), B, H, W, Cn, H, W, None)))
        bo = timeit(lambda: _lib.check(lib.b2f_warp_bhwd_backward(P(img), P(grid), P(go), None, P(gg), B, H, W, Cn, H, W, None)))
        px = B * H * W
        print("%dx%dx%dx%-3d %-6s fwd %6.1f us (%5.0f GB/s) | bwd %6.1f us (%5.0f GB/s) | bwd only-grid %6.1f us"
              % (B, H, W, Cn, name, f, 4 * px * (2 * Cn + 2) / f / 1e3, bw, 4 * px * (3 * Cn + 4) / bw / 1e3, bo))
