"""Time the whole-network forward (back2future_b200.pwc.PWCNet) and each convolution shape of it on the GPU.
usage: python tools/time_pwc.py [--B 1] [--H 448] [--W 1024] [--convs] [--no-image-warps]"""
import argparse
import ctypes as C
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from back2future_b200 import _lib, pwc


def time_it(fn, iters=20, warm=3):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--H", type=int, default=448)
    ap.add_argument("--W", type=int, default=1024)
    ap.add_argument("--convs", action="store_true")
    ap.add_argument("--past-flow", action="store_true")
    ap.add_argument("--no-image-warps", action="store_true")
    ap.add_argument("--tc", action="store_true", help="decoders on the tensor cores")
    ap.add_argument("--once", action="store_true", help="two eager forwards and nothing else (for an ncu launch list)")
    ap.add_argument("--by-op", action="store_true", help="time every call of the plan on its own, grouped by entry point")
    a = ap.parse_args()
    lib = _lib.load()
    net = pwc.PWCNet(pwc.Opt(past_flow=a.past_flow), image_warps=not a.no_image_warps, tensor_cores=a.tc)
    x = torch.randn(a.B, 9, a.H, a.W, device="cuda")
    p = net.plan(a.B, a.H, a.W)
    p.x.copy_(x)
    res = {"B": a.B, "H": a.H, "W": a.W}
    if a.once:
        net.run(p, graph=False)
        torch.cuda.synchronize()
        net.run(p, graph=False)
        torch.cuda.synchronize()
        print("launches per forward:", p.n_launches)
        return
    res["eager_ms"] = time_it(lambda: net.run(p, graph=False))
    res["graph_ms"] = time_it(lambda: net.run(p, graph=True))
    res["launches"] = p.n_launches
    macs = 0
    rows = []
    for op in p.ops:
        if op[0] == "fork" or op[1] is not lib.b2f_conv3x3_forward:
            continue
        args = op[2]
        B, Cin, H, W, Cout, stride = args[8], args[9], args[10], args[11], args[12], args[13]
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        m = B * Cin * 9 * Cout * Ho * Wo
        macs += m
        rows.append((args, m, (B, Cin, H, W, Cout, stride)))
    res["conv_gmac"] = macs / 1e9
    res["triplets_per_s_graph"] = a.B / res["graph_ms"] * 1e3
    res["conv_tflops_if_all_time_were_conv"] = 2 * macs / res["graph_ms"] / 1e9
    print(json.dumps(res))
    if a.by_op:
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        names = {getattr(lib, n)._name if hasattr(getattr(lib, n), "_name") else n: n for n in _lib.SIGNATURES}
        fn_name = {}
        for n in _lib.SIGNATURES:
            fn_name[id(getattr(lib, n))] = n
        agg = {}
        for op in p.ops:
            if op[0] == "fork":
                continue
            fn, args = op[1], op[2]
            ms = time_it(lambda: fn(*args, st), iters=5, warm=1)
            key = fn_name.get(id(fn), str(fn))
            if key == "b2f_conv3x3_forward":
                key += " (Cout=%d, stride %d)" % (args[12], args[13]) if args[12] <= 2 or args[13] == 2 else " (trunk)" if args[9] in (16, 32, 64, 96, 128, 192) and args[12] == args[9] else ""
            t, n = agg.get(key, (0.0, 0))
            agg[key] = (t + ms, n + 1)
        tot = sum(t for t, _ in agg.values())
        for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print("%-52s %3d calls %8.3f ms %5.1f%%" % (k, n, t, 100 * t / tot))
        print("sum of calls %.3f ms (serialised; the plan overlaps three lanes)" % tot)
    if a.convs and a.tc:
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        tot = 0.0
        for op in p.ops:
            if op[0] == "fork" or op[1] is not lib.b2f_conv3x3_tc_forward:
                continue
            args = op[2]
            B, Cin, H, W, Cout = args[9], args[10], args[11], args[12], args[13]
            m = B * Cin * 9 * Cout * H * W
            ms = time_it(lambda: lib.b2f_conv3x3_tc_forward(*args, st), iters=10, warm=2)
            tot += ms
            print("tc conv B=%d Cin=%3d %3dx%-4d Cout=%3d  %8.3f ms  %6.1f TFLOP/s (fp32-equivalent)" % (B, Cin, H, W, Cout, ms, 2 * m / ms / 1e9))
        print("sum of tensor-core convs %.3f ms" % tot)
    if a.convs:
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        tot = 0.0
        for args, m, shp in rows:
            ms = time_it(lambda: lib.b2f_conv3x3_forward(*args, st), iters=10, warm=2)
            tot += ms
            print("conv B=%d Cin=%3d %3dx%-4d Cout=%3d s=%d  %8.3f ms  %6.1f TFLOP/s" % (*shp, ms, 2 * m / ms / 1e9))
        print("sum of convs %.3f ms" % tot)


if __name__ == "__main__":
    main()
