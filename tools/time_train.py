"""Per-entry-point breakdown of one training step's backward plan (back2future_b200.pwc.PWCNet._build_backward).
usage: python tools/time_train.py [--B 8] [--H 320] [--W 640] [--past-flow]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from back2future_b200 import _lib, pwc


def time_it(fn, iters=5, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--H", type=int, default=320)
ap.add_argument("--W", type=int, default=640)
ap.add_argument("--past-flow", action="store_true")
ap.add_argument("--tc", action="store_true", help="tensor-core forward + input gradients (train_planar)")
ap.add_argument("--no-side", action="store_true", help="weight gradients on the main stream")
ap.add_argument("--once", action="store_true", help="two eager forward + backward plans and nothing else (for an ncu launch list)")
ap.add_argument("--detail", action="store_true", help="list every FFMA convolution call of the backward plan")
a = ap.parse_args()
lib = _lib.load()
if a.no_side:
    pwc.SIDE_LANES = False
net = pwc.PWCNet(pwc.Opt(past_flow=a.past_flow), tensor_cores=a.tc, train_planar=a.tc)
x = torch.randn(a.B, 9, a.H, a.W, device="cuda")
out = net.forward(x, graph=False)
net.backward(x, [torch.randn_like(t) for t in out])
torch.cuda.synchronize()
p = net.plan(a.B, a.H, a.W)
if a.once:
    p.launch()
    p.launch_backward()
    torch.cuda.synchronize()
    print("launches per forward: %d, backward entries: %d" % (p.n_launches, len(p.bops)))
    sys.exit(0)
name_of = {id(getattr(lib, n)): n for n in _lib.SIGNATURES}
st0 = C.c_void_p(torch.cuda.current_stream().cuda_stream)
fwd = time_it(lambda: p.launch())
bwd = time_it(lambda: p.launch_backward())
print("forward %.3f ms, backward %.3f ms (eager)" % (fwd, bwd))
if a.detail:
    fagg = {}
    for op in p.ops:
        if op[0] == "fork":
            continue
        lane, fn, args = op
        ms = time_it(lambda: fn(*args, st0), iters=3, warm=1)
        k = name_of.get(id(fn), "?")
        if k == "b2f_conv3x3_forward":
            ints = [v for v in args if isinstance(v, int)]
            print("   fwd %-40s %8.3f ms  ints %s" % (k, ms, ints[-7:]))
        t, n = fagg.get(k, (0.0, 0))
        fagg[k] = (t + ms, n + 1)
    for k, (t, n) in sorted(fagg.items(), key=lambda kv: -kv[1][0]):
        print("fwd %-42s %3d calls %8.3f ms" % (k, n, t))
name = {id(getattr(lib, n)): n for n in _lib.SIGNATURES}
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
agg = {}
for op in p.bops:
    fn, args = op[0], op[1]
    ms = time_it(lambda: fn(*args, st), iters=3, warm=1)
    k = name.get(id(fn), "?")
    if k == "b2f_conv3x3_backward_weights":
        k += " (stride %d)" % args[11]
    if k == "b2f_conv3x3_backward_data":
        k += " (stride %d)" % args[13]
    t, n = agg.get(k, (0.0, 0))
    agg[k] = (t + ms, n + 1)
    if a.detail and k.startswith(("b2f_conv3x3_backward_weights", "b2f_conv3x3_backward_data", "b2f_conv3x3_tc_")):
        ints = [v for v in args if isinstance(v, int)]
        print("   %-44s %8.3f ms  ints %s" % (k, ms, ints[-7:]))
tot = sum(t for t, _ in agg.values())
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-46s %3d calls %8.3f ms %5.1f%%" % (k, n, t, 100 * t / tot))
print("sum %.3f ms" % tot)
