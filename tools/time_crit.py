#!/usr/bin/env python
"""Time the fused criterion kernels at the training pyramid (BASELINE configs 3/4: B=8, 320x640 .. 20x40)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from back2future_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
B = 8
P = lambda t: C.c_void_p(t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=15):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
loss = torch.zeros(1, dtype=torch.float64, device=dev)
tot = {}
for k in range(5):
    h, w = 320 >> k, 640 >> k
    flow, bflow = torch.randn(B, 2, h, w, device=dev) * 0.2, torch.randn(B, 2, h, w, device=dev) * 0.2
    occ = torch.softmax(torch.randn(B, 2, h, w, device=dev), 1).contiguous()
    w1, w2, tgt = (torch.rand(B, 3, h, w, device=dev) * 4.7 - 2.1 for _ in range(3))
    g2a, g2b, g3a, g3b = torch.empty_like(flow), torch.empty_like(flow), torch.empty_like(w1), torch.empty_like(w1)
    px = B * h * w
    res = []
    for gt in (0, 1):
        prm = _lib.ObParams(gt, 1, 0.05, 1.0, 0.0 if gt else 1.0, 1.0, 1.0, 20.0 / 2 ** k, gt, 0, 0)
        t = timeit(lambda: _lib.check(lib.b2f_ob_criterion(C.byref(prm), P(flow), P(bflow), P(occ), P(w1), P(w2), P(tgt), B, 3, h, w,
                                                           P(g2a), P(g3a), P(g3b), P(loss), None, None)))
        res.append(("OBGCC" if gt else "OBCC", t, 84 * px))
    for order in (1, 2):
        prm = _lib.SmoothParams(order, 1, 0.05, 20.0, 0, 1)
        t = timeit(lambda: _lib.check(lib.b2f_smoothness_criterion(C.byref(prm), P(flow), P(tgt), B, 2, 3, h, w, P(g2a), P(loss), None, None)))
        res.append(("Smooth%d" % order, t, 28 * px))
    t = timeit(lambda: _lib.check(lib.b2f_constvel_criterion(P(flow), P(bflow), B, 2, h, w, 1, P(g2a), P(g2b), P(loss), None, None)))
    res.append(("ConstVel", t, 32 * px))
    t = timeit(lambda: _lib.check(lib.b2f_occprior_criterion(P(occ), B, 2, h, w, 1.0, 0, P(g2a), P(loss), None, None)))
    res.append(("OccPrior", t, 16 * px))
    print("%3dx%-3d " % (h, w) + " | ".join("%s %6.1f us %5.0f GB/s" % (n, t, by / t / 1e3) for n, t, by in res))
