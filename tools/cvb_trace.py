#!/usr/bin/env python
"""Timeline of the cost-volume backward kernel (cvb::costvol_bwd_tma<.,32,2>) from clock64 stamps.

Needs the instrumented build:  tools/build_variants.sh trace "-DB2F_CVB_TRACE"
Run (GPU):                     B2F_LIB_PATH=build/variants/trace.so python tools/cvb_trace.py [level]

Per compute warp (lane 0): slot 0 start, 1 %smid, 2+2*iy slab iy (and the X rows it needs) arrived,
3+2*iy FMAs of slab iy done, 20 store issued, 21 end.  Producer warp (warp 4): 2 slab 0 requested, 20 X tile
requested, 2+iy slab iy requested.  clock64 is per SM; CTAs are grouped by %smid.
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from back2future_b200 import _lib

lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
raw.b2f_debug_cvb_trace.argtypes = [C.c_void_p]
raw.b2f_debug_cvb_trace.restype = C.c_int
dev = torch.device("cuda:0")
l = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Cn = {3: 32, 4: 64, 5: 96}[l]
B, h, w = 8, 448 >> (l - 1), 1024 >> (l - 1)
ref, frm = torch.randn(B, Cn, h, w, device=dev), torch.randn(B, Cn, h, w, device=dev)
gj = torch.randn(B, 162, h, w, device=dev)
gr, gf = torch.empty_like(ref), torch.empty_like(frm)
fp = _lib.ptr_array([ref.data_ptr(), frm.data_ptr()])
gp = _lib.ptr_array([gr.data_ptr(), gf.data_ptr()])
P = lambda t: C.c_void_p(t.data_ptr())
lib.b2f_debug_costvol_path(int(os.environ.get("B2F_PATH", "14")))
run = lambda: _lib.check(lib.b2f_costvol_backward(fp, 2, B, Cn, h, w, 9, 1, P(gj), gj.stride(0), gp, None))
for _ in range(3):
    run()
torch.cuda.synchronize()
ncta = ((w + 31) // 32) * ((h + 7) // 8) * B * ((Cn + 31) // 32) * 2
trace = torch.zeros(ncta * 8 * 32, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush.zero_()
assert raw.b2f_debug_cvb_trace(C.c_void_p(trace.data_ptr())) == 0
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record()
torch.cuda.synchronize()
raw.b2f_debug_cvb_trace(None)
print("level %d: kernel %.1f us (instrumented), %d CTAs" % (l, a.elapsed_time(b) * 1e3, ncta))
t = trace.cpu().numpy().reshape(ncta, 8, 32)
GHZ = 1.0  # report cycles
cw = t[:, :4, :]                      # compute warps
start = cw[:, :, 0].min(axis=1)
end = cw[:, :, 21].max(axis=1)
smid = cw[:, 0, 1]
resid = end - start
first = cw[:, :, 2] - cw[:, :, 0]                      # wait for slab 0 + first X half
waits = np.stack([cw[:, :, 2 + 2 * iy] - cw[:, :, 3 + 2 * (iy - 1)] for iy in range(1, 9)], axis=2)
work = np.stack([cw[:, :, 3 + 2 * iy] - cw[:, :, 2 + 2 * iy] for iy in range(9)], axis=2)
epi = cw[:, :, 21] - cw[:, :, 19]
med = lambda x: float(np.median(x))
print("cycles per CTA (median over CTAs; per-warp values averaged over the 4 compute warps)")
print("  residency                    %8.0f" % med(resid))
print("  wait for slab 0 + X          %8.0f" % med(first.mean(axis=1)))
print("  waits for slabs 1..8 (sum)   %8.0f   per slab: %s" % (med(waits.sum(axis=2).mean(axis=1)),
      " ".join("%5.0f" % med(waits[:, :, k].mean(axis=1)) for k in range(8))))
print("  copy + FMA, 9 slabs (sum)    %8.0f   per slab: %s" % (med(work.sum(axis=2).mean(axis=1)),
      " ".join("%5.0f" % med(work[:, :, k].mean(axis=1)) for k in range(9))))
print("  epilogue                     %8.0f" % med(epi.mean(axis=1)))
# by wave: CTAs sorted by start on their SM
order = {}
for i in np.argsort(start):
    order.setdefault(int(smid[i]), []).append(i)
nw = max(len(v) for v in order.values())
print("per SM: %d SMs, CTAs per SM %d..%d" % (len(order), min(len(v) for v in order.values()), nw))
span = [max(end[v]) - min(start[v]) for v in order.values()]
print("  SM busy span (first start -> last end): median %.0f  max %.0f cycles" % (np.median(span), max(span)))
# SM-level: fraction of the span in which at least one compute warp of the SM is inside copy+FMA, and both CTAs are
tot_any = tot_cnt = 0.0
for sm, v in order.items():
    ev = []
    for i in v:
        for wv in range(4):
            for iy in range(9):
                ev.append((cw[i, wv, 2 + 2 * iy], 1)); ev.append((cw[i, wv, 3 + 2 * iy], -1))
    ev.sort()
    cur = 0; last = ev[0][0]; any_t = 0; cnt_t = 0
    for tt, d in ev:
        if cur > 0:
            any_t += tt - last
        cnt_t += cur * (tt - last)
        cur += d; last = tt
    sp = max(end[v]) - min(start[v])
    tot_any += any_t / sp; tot_cnt += cnt_t / sp
print("  share of the SM span with >= 1 warp in copy+FMA: %.2f ; mean warps in copy+FMA: %.2f (of 8)" %
      (tot_any / len(order), tot_cnt / len(order)))
# one SM as text
sm0 = sorted(order)[len(order) // 2]
print("SM %d timeline (cycles from the SM's first CTA start): cta start first_slab end | per-slab wait/work of warp 0" % sm0)
t0 = min(start[order[sm0]])
for i in order[sm0]:
    ws = " ".join("%d/%d" % ((cw[i, 0, 2 + 2 * iy] - (cw[i, 0, 3 + 2 * (iy - 1)] if iy else cw[i, 0, 0])),
                              cw[i, 0, 3 + 2 * iy] - cw[i, 0, 2 + 2 * iy]) for iy in range(9))
    print("  cta %5d  %7d %7d %7d | %s" % (i, start[i] - t0, cw[i, 0, 2] - t0, end[i] - t0, ws))
