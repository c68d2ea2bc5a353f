#!/usr/bin/env python
"""Profile target: launches each hot kernel of BASELINE config 2 a few times at its real size.
Usage (under gpurun): ncu ... python tools/prof_target.py [cv3|cv4|warp|all] [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from back2future_b200 import _lib

lib = _lib.load()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
P = lambda t: C.c_void_p(t.data_ptr())
B = 8


def cv(level_c, h, w):
    ref, frm = (torch.randn((B, level_c, h, w), device=dev, generator=g) for _ in range(2))
    joined = torch.empty((B, 162, h, w), device=dev)
    gj = torch.randn((B, 162, h, w), device=dev, generator=g)
    gr, gf = torch.empty_like(ref), torch.empty_like(frm)
    fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()])
    gp = _lib.ptr_array([gr.data_ptr(), gf.data_ptr()])
    for _ in range(reps):
        for fwd, half in ((1, 0), (0, 1)):
            _lib.check(lib.b2f_costvol_forward(fp, 2, B, level_c, h, w, 9, fwd, P(joined[:, 81 * half:]), joined.stride(0), None))
            _lib.check(lib.b2f_costvol_backward(fp, 2, B, level_c, h, w, 9, fwd, P(gj[:, 81 * half:]), gj.stride(0), gp, None))
    torch.cuda.synchronize()


def warp(Cn, h, w):
    img = torch.randn((B, h, w, Cn), device=dev, generator=g)
    grid = torch.randn((B, h, w, 2), device=dev, generator=g) * 4
    go = torch.randn((B, h, w, Cn), device=dev, generator=g)
    out, gi, gg = torch.empty_like(img), torch.zeros_like(img), torch.empty_like(grid)
    for _ in range(reps):
        _lib.check(lib.b2f_warp_bhwd_forward(P(img), P(grid), P(out), B, h, w, Cn, h, w, None))
        _lib.check(lib.b2f_warp_bhwd_backward(P(img), P(grid), P(go), P(gi), P(gg), B, h, w, Cn, h, w, None))
    torch.cuda.synchronize()


def warp_unit(Cn, h, w, scale):
    """fused warpingUnit (row N2), smooth flow in network units"""
    img = torch.randn((B, Cn, h, w), device=dev, generator=g)
    go = torch.randn((B, Cn, h, w), device=dev, generator=g)
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
    k = w / 1024.0
    flow = (torch.stack([6 * k * torch.sin(xs / (90 * k) + ys / (70 * k)) + 3 * k, 5 * k * torch.cos(xs / (60 * k) - ys / (110 * k))], 0)[None]
            + 0.05 * torch.randn((B, 2, h, w), device=dev, generator=g)).contiguous() / scale
    out, gi, gf = torch.empty_like(img), torch.zeros_like(img), torch.empty_like(flow)
    for _ in range(reps):
        _lib.check(lib.b2f_warp_bdhw_forward(P(img), P(flow), scale, P(out), B, Cn, h, w, None))
        _lib.check(lib.b2f_warp_bdhw_backward(P(img), P(flow), scale, P(go), P(gi), P(gf), B, Cn, h, w, None))
    torch.cuda.synchronize()


if what == "wu":
    warp_unit(32, 112, 256, 5.0)
    warp_unit(3, 448, 1024, 20.0)
if what in ("cv3", "all"):
    cv(32, 112, 256)
if what in ("cv4", "all"):
    cv(64, 56, 128)
if what in ("cv5", "all"):
    cv(96, 28, 64)
if what in ("warp", "all"):
    warp(3, 448, 1024)
    warp(32, 112, 256)
