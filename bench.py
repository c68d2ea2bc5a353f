#!/usr/bin/env python
"""bench.py -- frame-triplets/sec of the Back2Future hot path at 1024x448 on N B200s.

Workload (BASELINE.json configs[1]): CostVolMulti (past + future) and BilinearSamplerBHWD forward +
backward over the full pyramid of the Ours-Hard architecture, batch 8 triplets per GPU, 1024x448,
synthetic N(0,1) features, flow ~ N(0, 4 px):
  * cost volume, levels 3..7 (C = 32,64,96,128,192; 112x256 .. 7x16), both directions, forward into
    the 162-channel joined buffer and backward from a narrow of the 162-channel gradient;
  * feature warps at levels 6..3 (C = 128,96,64,32) x 2 frames, forward + backward (image + flow grad);
  * image warps (C = 3) at 28x64 .. 448x1024 x 2 frames, forward + backward.
A "step" is one pass of all of that over one batch.  Triplets are independent, so ranks shard them
with no data-path collective (weak scaling).

  value : device-resident throughput (inputs already in HBM), C-ABI calls on one stream.
  e2e   : the same pass through the reference-facing module API (back2future_b200.nn) with HOST
          buffers: every step copies its inputs from pinned host memory and reads every result back.
  --impl reference : the CPU restatement of the reference algorithm (oracle/c, OpenMP, all host
          threads) on a bounded sample of the same workload.  Torch7 itself cannot be installed.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H_FULL, W_FULL = 448, 1024
LEVEL_C = {3: 32, 4: 64, 5: 96, 6: 128, 7: 192}
BATCH = 8
METRIC = "frame_triplets_per_sec_1024x448"
UNIT = "triplets/s"
WORKLOAD = ("configs[1]: CostVolMulti + BilinearSamplerBHWD fwd/bwd microbench, full pyramid "
            "(levels 3-7, both directions; feature warps L6-L3; image warps 28x64..448x1024), "
            "batch 8 per GPU, 1024x448 synthetic")


def level_hw(l):
    return H_FULL >> (l - 1), W_FULL >> (l - 1)


# ----------------------------------------------------------------------------------------
# clocks (recipe: sample DURING the timed region)
# ----------------------------------------------------------------------------------------

class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------
# the workload on one GPU
# ----------------------------------------------------------------------------------------

class Op:
    __slots__ = ("name", "kind", "call", "bytes", "flops", "zero")

    def __init__(self, name, kind, call, nbytes, flops=0, zero=None):
        self.name, self.kind, self.call, self.bytes, self.flops, self.zero = name, kind, call, nbytes, flops, zero


class _Zero:
    """A pending b2f_zero_async(ptr, bytes, stream) call."""

    def __init__(self, lib, arg):
        self.lib, self.arg = lib, arg

    def zero_on(self, stream_handle):
        return self.lib.b2f_zero_async(self.arg[0], self.arg[1], stream_handle)


class Workload:
    """Device buffers + the ordered list of C-ABI calls of one step."""

    def __init__(self, torch, lib, dev, B=BATCH, seed=2, full_hw=None):
        from back2future_b200 import _lib
        self.torch, self.lib, self.dev, self.B = torch, lib, dev, B
        g = torch.Generator(device=dev).manual_seed(seed)
        HF, WF = full_hw or (H_FULL, W_FULL)      # network input size (a multiple of 64, back2future.lua:57-71)

        def level_hw(l):
            return HF >> (l - 1), WF >> (l - 1)

        def randn(*shape, scale=1.0):
            t = torch.randn(shape, device=dev, generator=g, dtype=torch.float32)
            return t * scale if scale != 1.0 else t

        def empty(*shape):
            return torch.empty(shape, device=dev, dtype=torch.float32)

        self.grids = []         # the flow fields of the warps (refilled by set_flow for the flow-variant timings)
        self.keep = []          # every tensor the ops reference
        self.inputs = []        # (name, tensor) copied H2D per e2e step
        self.outputs = []       # (name, tensor) copied D2H per e2e step
        self.fwd_ops, self.bwd_ops = [], []
        st = None               # stream handle filled in by bind()
        self._mk = []           # deferred op constructors (need the stream)
        P = lambda t: C.c_void_p(t.data_ptr())

        # ---- cost volumes -------------------------------------------------------------
        for l in (7, 6, 5, 4, 3):
            Cn = LEVEL_C[l]
            h, w = level_hw(l)
            ref, past, fut = randn(B, Cn, h, w), randn(B, Cn, h, w), randn(B, Cn, h, w)
            joined = empty(B, 162, h, w)
            gjoined = randn(B, 162, h, w)
            grads = [empty(B, Cn, h, w) for _ in range(4)]
            self.keep += [ref, past, fut, joined, gjoined] + grads
            self.inputs += [("cv%d.ref" % l, ref), ("cv%d.past" % l, past), ("cv%d.fut" % l, fut),
                            ("cv%d.gradJoined" % l, gjoined)]
            self.outputs += [("cv%d.joined" % l, joined)] + [("cv%d.grad%d" % (l, i), t) for i, t in enumerate(grads)]
            fb = 4 * B * h * w * (2 * Cn + 81)
            bb = 4 * B * h * w * (81 + 4 * Cn)
            fl = 2 * Cn * 81 * B * h * w
            for d, (frame, fwd, half) in enumerate(((fut, 1, 0), (past, 0, 1))):
                fptr = _lib.ptr_array([ref.data_ptr(), frame.data_ptr()])
                out = joined[:, 81 * half:81 * (half + 1)]
                go = gjoined[:, 81 * half:81 * (half + 1)]
                gptr = _lib.ptr_array([grads[2 * d].data_ptr(), grads[2 * d + 1].data_ptr()])
                self.keep += [fptr, gptr]
                self._mk.append(("f", "costvol_fwd L%d %s" % (l, "fut" if fwd else "past"), "costvol_fwd_L%d" % l,
                                 lambda s, a=(fptr, 2, B, Cn, h, w, 9, fwd, P(out), joined.stride(0)):
                                 (lambda: lib.b2f_costvol_forward(*a, s)), fb, fl, None, ("cv", l, d)))
                self._mk.append(("b", "costvol_bwd L%d %s" % (l, "fut" if fwd else "past"), "costvol_bwd_L%d" % l,
                                 lambda s, a=(fptr, 2, B, Cn, h, w, 9, fwd, P(go), gjoined.stride(0), gptr):
                                 (lambda: lib.b2f_costvol_backward(*a, s)), bb, 2 * fl, None, ("cv", l, d)))

        # ---- warps ---------------------------------------------------------------------
        # feature warp "L l" produces the warped features the level-l cost volume consumes (pwc.lua:402-408); image
        # warp k belongs to pyramid level l = k + 3: it warps the image at the resolution of level l - 2 with the
        # skip-upsampled flow of level l (pwc.lua:441-446)
        warp_cfgs = [("feat L%d" % l, LEVEL_C[l], level_hw(l), ("fw", l)) for l in (6, 5, 4, 3)]
        warp_cfgs += [("img %dx%d" % (HF >> k, WF >> k), 3, (HF >> k, WF >> k), ("iw", k + 3))
                      for k in (4, 3, 2, 1, 0)]
        for name, Cn, (h, w), wtag in warp_cfgs:
            for d, fr in ((1, "past"), (0, "fut")):
                img, grid = randn(B, h, w, Cn), randn(B, h, w, 2, scale=4.0)
                self.grids.append(grid)
                out, go = empty(B, h, w, Cn), randn(B, h, w, Cn)
                gimg, ggrid = empty(B, h, w, Cn), empty(B, h, w, 2)
                self.keep += [img, grid, out, go, gimg, ggrid]
                tag = "warp %s %s" % (name, fr)
                self.inputs += [(tag + ".img", img), (tag + ".grid", grid), (tag + ".gradOut", go)]
                self.outputs += [(tag + ".out", out), (tag + ".gradImg", gimg), (tag + ".gradGrid", ggrid)]
                fb = 4 * B * h * w * (2 * Cn + 2)
                bb = 4 * B * h * w * (3 * Cn + 4)
                kind = "warp_%s" % name.replace(" ", "_")
                self._mk.append(("f", tag + " fwd", kind + "_fwd",
                                 lambda s, a=(P(img), P(grid), P(out), B, h, w, Cn, h, w):
                                 (lambda: lib.b2f_warp_bhwd_forward(*a, s)), fb, 0, None, wtag + (d,)))
                self._mk.append(("b", tag + " bwd", kind + "_bwd",
                                 lambda s, a=(P(img), P(grid), P(go), P(gimg), P(ggrid), B, h, w, Cn, h, w):
                                 (lambda: lib.b2f_warp_bhwd_backward(*a, s)), bb, 0,
                                 (P(gimg), gimg.numel() * 4), wtag + (d,)))

    @staticmethod
    def network_order():
        """(direction, tag) of every call of one step in the order the network's dataflow imposes
        (models/pwc.lua:237-456): coarse to fine, per level the two cost volumes, then the level's image warps
        (leaves of the loss) and the feature warps that feed the next finer level; backward is the mirror image."""
        fwd = []
        for l in (7, 6, 5, 4, 3):
            fwd += [("cv", l, 0), ("cv", l, 1), ("iw", l, 0), ("iw", l, 1)]
            if l > 3:
                fwd += [("fw", l - 1, 0), ("fw", l - 1, 1)]
        return [("f", t) for t in fwd] + [("b", t) for t in fwd[::-1]]

    def bind(self, stream_handle):
        s = C.c_void_p(stream_handle)
        lib = self.lib
        self.fwd_ops, self.bwd_ops = [], []
        by_tag = {}
        for which, name, kind, mk, nbytes, flops, zero, tag in self._mk:
            z = None
            if zero is not None:
                z = (lambda a=zero: lib.b2f_zero_async(a[0], a[1], s))
            op = Op(name, kind, mk(s), nbytes, flops, z)
            (self.fwd_ops if which == "f" else self.bwd_ops).append(op)
            by_tag[(which, tag)] = op
        self.ops = [by_tag[k] for k in self.network_order()]
        assert len(self.ops) == len(self._mk)

    def step(self, marks=None):
        """One pass: forward coarse-to-fine, then backward in reverse.  `marks` maps an op kind to a
        list receiving (start_event, end_event) pairs for the live roofline measurement."""
        torch = self.torch
        for op in self.ops:
            if op.zero is not None:
                rc = op.zero()
                if rc:
                    raise RuntimeError("b2f_zero_async failed: %d" % rc)
            if marks is not None and op.kind in marks:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = op.call()
                e1.record()
                marks[op.kind].append((e0, e1, op))
            else:
                rc = op.call()
            if rc:
                raise RuntimeError("%s failed: status %d: %s" % (op.name, rc, self.lib.b2f_last_error().decode()))

    def total_bytes(self):
        return sum(op.bytes for op in self.ops)

    def set_flow(self, kind, seed=3):
        """Refill every warp's flow field in place: "iid4" / "iid0.5" = i.i.d. N(0, sigma px) (sigma = 4 is
        the headline workload, 0.5 its second run), "smooth" = a low-frequency field of +-6 px (scaled with the
        level's width) plus 0.05 px noise -- what the decoder's bilinearly up-sampled flow looks like."""
        torch = self.torch
        g = torch.Generator(device=self.dev).manual_seed(seed)
        for grid in self.grids:
            B, h, w, _ = grid.shape
            if kind.startswith("iid"):
                grid.copy_(torch.randn(grid.shape, device=self.dev, generator=g) * float(kind[3:]))
            else:
                ys, xs = torch.meshgrid(torch.arange(h, device=self.dev, dtype=torch.float32),
                                        torch.arange(w, device=self.dev, dtype=torch.float32), indexing="ij")
                k = w / 1024.0
                f = torch.stack([6 * k * torch.sin(xs / (90 * k) + ys / (70 * k)) + 3 * k,
                                 5 * k * torch.cos(xs / (60 * k) - ys / (110 * k))], -1)
                grid.copy_(f[None] + 0.05 * torch.randn(grid.shape, device=self.dev, generator=g))

    # -- CUDA graph of one step -------------------------------------------------------------
    def capture(self, streams, forward_only=False):
        """Capture one step as a CUDA graph whose edges are the network's own data dependencies
        (models/pwc.lua:237-456), over six streams (the first one is the capture stream):

          forward, per level l = 7..3:   [CV_l future || CV_l past] -> join (the level's decoder)
                                          -> image warps of level l (side streams; leaves of the loss: nothing in the
                                             forward waits for them) and, for l > 3, [feature warp future || past]
                                             -> join -> CV_{l-1};
          backward, l = 3..7:            image-warp backward of level l (needs only the loss gradient, so all five
                                          levels start when the backward starts; level 3 on the main streams) must be
                                          done before [CV_l backward future || past] -> [feature-warp L l backward].

        So the L1/L2-bound C = 3 image warps of a level run under the shared-memory/FMA-bound cost volumes of the next
        level, as a graph executor of the real network would schedule them.  The gradImg zero-fills (b2f_zero_async)
        are issued at the start of the step on their own stream and joined before the first scatter."""
        torch, lib = self.torch, self.lib
        assert len(streams) >= 6
        M, S, I1, I2, Z = streams[0], streams[1], streams[2], streams[3], streams[5]
        H = {st: C.c_void_p(st.cuda_stream) for st in streams}
        ops = {}
        for which, name, kind, mk, nbytes, flops, zero, tag in self._mk:
            ops[(which, tag)] = (name, mk, zero)
        sched = []

        def call(which, tag, st):
            name, mk, _ = ops[(which, tag)]
            rc = mk(H[st])()
            if rc:
                raise RuntimeError("%s failed during capture: %d: %s" % (name, rc, lib.b2f_last_error().decode()))
            sched.append((name, streams.index(st)))

        def pair(which, kind, l):
            S.wait_stream(M)
            call(which, (kind, l, 0), M)
            call(which, (kind, l, 1), S)
            M.wait_stream(S)

        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=M):
            Z.wait_stream(M)
            for which, name, kind, mk, nbytes, flops, zero, tag in self._mk:
                if forward_only:
                    break
                if zero is not None and lib.b2f_zero_async(zero[0], zero[1], H[Z]):
                    raise RuntimeError("b2f_zero_async failed during capture")
            # ---- forward
            for l in (7, 6, 5, 4, 3):
                pair("f", "cv", l)
                I1.wait_stream(M)
                I2.wait_stream(M)
                call("f", ("iw", l, 0), I1)
                call("f", ("iw", l, 1), I2)
                if l > 3:
                    pair("f", "fw", l - 1)
            M.wait_stream(I1)
            M.wait_stream(I2)
            M.wait_stream(Z)          # every scatter target is clean before the first backward kernel
            if forward_only:          # inference (back2future.lua:74 model:forward): the forward half is the step
                self.graph_schedule = sched
                return graph
            # ---- backward
            I1.wait_stream(M)
            I2.wait_stream(M)
            pair("b", "iw", 3)
            done = {}
            for l in (4, 5, 6, 7):
                call("b", ("iw", l, 0), I1)
                call("b", ("iw", l, 1), I2)
                e1, e2 = torch.cuda.Event(), torch.cuda.Event()
                e1.record(I1)
                e2.record(I2)
                done[l] = (e1, e2)
            for l in (3, 4, 5, 6, 7):
                if l > 3:
                    M.wait_event(done[l][0])
                    M.wait_event(done[l][1])
                pair("b", "cv", l)
                if l < 7:
                    pair("b", "fw", l)
            M.wait_stream(I1)
            M.wait_stream(I2)
        self.graph_schedule = sched
        assert len(sched) == len(self._mk)
        return graph


def breakdown(torch, wl, iters=10):
    """Per-kernel device times (separate pass, events around every call; not part of `value`)."""
    acc = {}
    for _ in range(iters):
        evs = []
        for op in wl.ops:
            if op.zero is not None:
                op.zero()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op.call()
            e1.record()
            evs.append((op, e0, e1))
        torch.cuda.synchronize()
        for op, e0, e1 in evs:
            acc.setdefault(op.name, [op, []])[1].append(e0.elapsed_time(e1))
    rows = []
    for name, (op, ts) in acc.items():
        ts = sorted(ts)
        ms = ts[len(ts) // 2]
        rows.append({"kernel": name, "ms": round(ms, 5), "alg_MB": round(op.bytes / 1e6, 2),
                     "GBps": round(op.bytes / ms / 1e6, 1) if ms > 0 else None,
                     "GFLOPs": round(op.flops / ms / 1e6, 1) if op.flops and ms > 0 else None})
    return rows


# ----------------------------------------------------------------------------------------
# inference shapes (BASELINE configs[0], [4]): hot-path-only, forward only, one triplet at a time
# ----------------------------------------------------------------------------------------

def time_inference(torch, lib, dev, reps=200):
    """The hot-path calls of ONE inference forward (back2future.lua:74: both cost volumes of levels 7..3, the feature
    warps, the image warps; no backward) at the sizes `computeFlow` feeds the network for BASELINE configs[0]
    (1242x375 -> 1216x320) and configs[4] (1024x436 -> 1024x384), B = 1, synthetic features, as one CUDA-graph replay
    per triplet.  The conv trunk is not part of this library, so this is the hot path's share of a
    triplet, not whole-network triplets/s."""
    out = {}
    for key, hw in (("config0_1216x320", (320, 1216)), ("config4_1024x384", (384, 1024))):
        wl = Workload(torch, lib, dev, B=1, full_hw=hw)
        wl.bind(torch.cuda.current_stream().cuda_stream)
        nf = len(wl.ops) // 2
        for op in wl.ops[:nf]:           # eager forward once: kernel attributes, tensor maps
            if op.call():
                raise RuntimeError("%s failed: %s" % (op.name, lib.b2f_last_error().decode()))
        torch.cuda.synchronize()
        graph = wl.capture([torch.cuda.Stream(device=dev) for _ in range(6)], forward_only=True)
        res = {"launches_per_triplet": nf, "alg_bytes_per_triplet": sum(op.bytes for op in wl.ops[:nf])}
        for kind in ("smooth", "iid4"):
            wl.set_flow(kind)
            for _ in range(5):
                graph.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                graph.replay()
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / reps
            res["us_per_triplet_%s_flow" % kind] = round(us, 1)
        res["hot_path_triplets_per_s"] = round(1e6 / res["us_per_triplet_smooth_flow"], 1)
        out[key] = res
        del wl, graph
    out["note"] = ("forward-only hot path of one triplet (B = 1) as one CUDA-graph replay; whole-network inference "
                   "(conv trunk included) is the `whole_network` block")
    return out


# ----------------------------------------------------------------------------------------
# whole network: Ours-Hard inference, device-resident and end to end from host images
# ----------------------------------------------------------------------------------------

def time_network(torch, dev, world, rank, dist, B=BATCH, steps=10, cpu=True):
    """Ours-Hard forward (back2future_b200.pwc.PWCNet: conv trunk + cost volumes + warps + decoders as one CUDA-graph
    replay) at 1024 x 448.  `device`: inputs resident, B triplets per replay.  `e2e`: every step uploads the B x 9 x H x W
    normalised frames from pinned host memory and reads flow + occlusion maps of the finest level back (what computeFlow
    moves, back2future.lua:73-92); H2D, compute and D2H run on three streams, two input / output buffers deep.
    `b1_*`: one triplet per replay (the reference's computeFlow call pattern) at the three BASELINE sizes.
    `cpu`: oracle/pwc_oracle.py (numpy float32, BLAS threads) on one triplet -- the checker, timed beside it."""
    from back2future_b200 import pwc
    import numpy as np
    out = {"model": "Ours-Hard (7 193 316 parameters, random init)", "batch_per_gpu": B}

    def timed(fn, n, sync_streams=()):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for st in sync_streams:
            st.wait_event(a)
        for _ in range(n):
            fn()
        for st in sync_streams:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream().wait_event(ev)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    out["decoders"] = ("tcgen05 implicit GEMM, three-pass TF32 split over channel-minor (hi, lo) activations "
                       "(b2f_conv3x3_tc_forward), the 2-channel heads included; so are the stride-1 pyramid layers with "
                       ">= 64 channels; the other pyramid layers on the FFMA2 kernel")
    net = pwc.PWCNet(pwc.Opt(), device=dev, image_warps=True, tensor_cores=True)
    p = net.plan(B, H_FULL, W_FULL)
    p.x.copy_(torch.randn(p.x.shape, device=dev))
    ms = timed(lambda: net.run(p), steps)
    out["device"] = {"triplets_per_s": round(world * B / ms * 1e3, 1), "ms_per_step": round(ms, 3),
                     "launches_per_step": p.n_launches, "outputs": "full table incl. the ten warped frames"}
    # two plans (two sets of buffers, two graphs) replayed on two streams: one batch's coarse levels -- launches that
    # cannot fill the machine, chained by the coarse-to-fine dependency -- run under the other batch's fine levels
    pb = net.plan(B, H_FULL, W_FULL, slot=1)
    pb.x.copy_(torch.randn(pb.x.shape, device=dev))
    sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for q, s_ in ((p, sa), (pb, sb)):          # capture each plan's graph
        with torch.cuda.stream(s_):
            net.run(q)
    torch.cuda.synchronize()

    def two():
        with torch.cuda.stream(sa):
            net.run(p)
        with torch.cuda.stream(sb):
            net.run(pb)

    ms2 = timed(two, steps, (sa, sb)) / 2
    out["device_two_streams"] = {"triplets_per_s": round(world * B / ms2 * 1e3, 1), "ms_per_batch": round(ms2, 3),
                                 "what": "two batches of %d in flight: two plans (buffer sets, graphs) on two streams" % B}
    del p, pb, net
    # the same network with every convolution on the fp32 FMA pipe (the training path's forward)
    net = pwc.PWCNet(pwc.Opt(), device=dev, image_warps=True, tensor_cores=False)
    p = net.plan(B, H_FULL, W_FULL)
    p.x.copy_(torch.randn(p.x.shape, device=dev))
    ms_f = timed(lambda: net.run(p), steps)
    out["device_ffma_only"] = {"triplets_per_s": round(world * B / ms_f * 1e3, 1), "ms_per_step": round(ms_f, 3)}
    # end to end: flow + occlusion only (computeFlow never reads the warped frames)
    net2 = pwc.PWCNet(pwc.Opt(), device=dev, image_warps=False, tensor_cores=True)
    p2 = net2.plan(B, H_FULL, W_FULL)
    # two plans on two compute streams (as `device_two_streams`); every step uploads its frames straight into its plan's
    # input buffer and reads flow + occlusion straight from the plan's outputs
    plans = [p2, net2.plan(B, H_FULL, W_FULL, slot=1)]
    hin = [torch.randn(B, 9, H_FULL, W_FULL).pin_memory() for _ in range(2)]
    hout = [[torch.empty(B, 2, H_FULL, W_FULL).pin_memory() for _ in range(2)] for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_cmp = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    ev_cmp, ev_out = [None, None], [None, None]
    state = {"i": 0}

    def e2e_step():
        i = state["i"] & 1
        state["i"] += 1
        q, sc = plans[i], s_cmp[i]
        with torch.cuda.stream(s_in):
            if ev_cmp[i] is not None:
                s_in.wait_event(ev_cmp[i])             # this plan's previous batch has been computed
            q.x.copy_(hin[i], non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record()
        with torch.cuda.stream(sc):
            sc.wait_event(ev_in)
            if ev_out[i] is not None:
                sc.wait_event(ev_out[i])               # ... and its results have been read back
            net2.run(q)
            ev = torch.cuda.Event()
            ev.record()
            ev_cmp[i] = ev
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev)
            hout[i][0].copy_(q.output[0], non_blocking=True)
            hout[i][1].copy_(q.output[1], non_blocking=True)
            eo = torch.cuda.Event()
            eo.record()
            ev_out[i] = eo

    ems = timed(e2e_step, steps, (s_in, s_out, s_cmp[0], s_cmp[1]))
    out["e2e"] = {"triplets_per_s": round(world * B / ems * 1e3, 1), "ms_per_step": round(ems, 3),
                  "h2d_bytes_per_step": B * 9 * H_FULL * W_FULL * 4, "d2h_bytes_per_step": 2 * B * 2 * H_FULL * W_FULL * 4,
                  "api": "back2future_b200.pwc.PWCNet.run on host frames (the device half of computeFlow), two batches in "
                         "flight (two plans on two streams)"}
    din = dout = None
    del net, p, hin, din, dout, hout
    if rank == 0:
        # the dominant kernel of the whole-network forward against the tensor-core roofline: one level-3 decoder layer
        # (128 -> 128 channels at 112 x 256 x B), CUDA events around 20 launches on resident operands
        import ctypes as C
        from back2future_b200 import _lib
        lib = _lib.load()
        Hc, Wc, Cc = H_FULL // 4, W_FULL // 4, 128
        xh = torch.randn(B, Hc, Wc, Cc, device=dev)
        xl = torch.randn(B, Hc, Wc, Cc, device=dev) * 1e-4
        nw = int(lib.b2f_conv3x3_tc_packed_floats(Cc, Cc))
        wh, wl = torch.randn(nw, device=dev) * 0.03, torch.randn(nw, device=dev) * 1e-5
        oh, ol = torch.empty_like(xh), torch.empty_like(xh)
        vp = lambda t: C.c_void_p(t.data_ptr())
        stc = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        call = lambda: lib.b2f_conv3x3_tc_forward(vp(xh), vp(xl), vp(wh), vp(wl), None, vp(oh), vp(ol), None, 0, B, Cc, Hc, Wc,
                                                  Cc, 0.2, stc)
        for _ in range(3):
            _lib.check(call())
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(20):
            call()
        b_.record()
        torch.cuda.synchronize()
        cms = a_.elapsed_time(b_) / 20
        flop = 2.0 * B * Hc * Wc * Cc * Cc * 9
        peak_bf16 = None
        try:
            peak_bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            pass
        peak_tf32 = (peak_bf16 or 2250.0) / 2
        out["roofline"] = {"bound": "tensor", "kernel": "conv3x3_tc_kernel<128> (decoder layer 128 -> 128, level 3)",
                           "achieved": round(3 * flop / cms / 1e9, 1), "peak": round(peak_tf32, 1), "unit": "TFLOP/s",
                           "frac": round(3 * flop / cms / 1e9 / peak_tf32, 4), "traffic": None,
                           "avg_launch_ms": round(cms, 4), "fp32_equivalent_tflops": round(flop / cms / 1e9, 1),
                           "flops_per_launch": 3 * flop,
                           "note": "achieved counts the three TF32 passes of the split (the MMAs executed); peak = half the "
                                   "measured dense bf16 rate of MEASURED_PEAKS.json (%s)" % ("measured" if peak_bf16 else "nominal fallback")}
        del xh, xl, wh, wl, oh, ol
    if rank == 0 and world == 1:
        for key, (h, w) in (("b1_1024x448", (H_FULL, W_FULL)), ("b1_config0_1216x320", (320, 1216)),
                            ("b1_config4_1024x384", (384, 1024))):
            q = net2.plan(1, h, w)
            q.x.copy_(torch.randn(q.x.shape, device=dev))
            m1 = timed(lambda: net2.run(q), 20)
            out[key] = {"ms_per_triplet": round(m1, 3), "triplets_per_s": round(1e3 / m1, 1)}
        # parity of this very model at a size the oracle finishes in a second
        from oracle import pwc_oracle as po, b2f_oracle as o
        params = pwc.PWCNet.random_params(net2.opt, 2)
        x = np.random.default_rng(5).uniform(-2.1, 2.6, (1, 9, 64, 128)).astype(np.float32)
        got = net2.forward(torch.from_numpy(x).to(dev))
        torch.cuda.synchronize()
        ref = po.pwc_forward(params, x, po.Opt())
        per = 4
        errs = [o.rel_err(got[2 * k].cpu().numpy(), ref[per * k]) for k in range(5)] + \
               [o.rel_err(got[2 * k + 1].cpu().numpy(), ref[per * k + 1]) for k in range(5)]
        out["parity"] = {"max_rel_err": float("%.3g" % max(errs)), "tol": 1e-4, "passed": max(errs) < 1e-4,
                         "what": "flow + occlusion of all 5 levels, 1 x 9 x 64 x 128, vs oracle/pwc_oracle.py (float64)"}
        if cpu:
            x = np.random.default_rng(6).uniform(-2.1, 2.6, (1, 9, H_FULL, W_FULL)).astype(np.float32)
            t0 = time.perf_counter()
            po.pwc_forward(params, x, po.Opt(), dtype=np.float32)
            dt = time.perf_counter() - t0
            out["cpu"] = {"triplets_per_s": round(1.0 / dt, 3), "s_per_triplet": round(dt, 2), "kind": "port",
                          "cores": len(os.sched_getaffinity(0)),
                          "sample": "1 triplet at 1024x448 through oracle/pwc_oracle.py (numpy float32 / BLAS)"}
    return out


# ----------------------------------------------------------------------------------------
# training steps (BASELINE configs[2], [3]): forward + criterions + backward + all-reduce + Adam
# ----------------------------------------------------------------------------------------

def time_training(torch, dev, world, rank, dist, B=BATCH, steps=6):
    """One optimisation step of train.lua:196-496 through back2future_b200.train.Trainer at the reference's training
    shape (8 samples per GPU, 9 x 320 x 640), inputs uploaded from pinned host memory every step (inside the timed
    region; the upload of step k + 1 is issued by step k and runs under its kernels: two host batches alternate),
    losses read back every step (cutorch.synchronize, train.lua:498).  With N > 1 the flat gradient is
    all-reduced through libb2f_comm.so in 11 buckets started from events inside the backward; `allreduce_exposed_ms`
    is the step time minus the same step without the collective."""
    from back2future_b200 import pwc, train, comm as bcomm
    H, W = 320, 640
    out = {"shape": "%d x 9 x %d x %d per GPU" % (B, H, W)}
    cm = bcomm.Communicator.from_env() if world > 1 else None
    hin = torch.empty(B, 9, H, W).uniform_(-2.1, 2.6).pin_memory()
    hin2 = hin.clone().pin_memory()
    def timed(net, topt, c):
        tr = train.Trainer(net, topt, comm=c)
        # every step uploads its own inputs from pinned host memory; the upload of step k + 1 is started inside step k
        # (train_batch's `prefetch`: double buffering), two host batches alternate
        for _ in range(2):
            losses = tr.train_batch(hin, prefetch=hin2)
            losses = tr.train_batch(hin2, prefetch=hin)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(steps):
            cur_in, nxt = (hin, hin2) if k % 2 == 0 else (hin2, hin)
            losses = tr.train_batch(cur_in, prefetch=nxt)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        del tr
        return ms, losses

    for key, past_flow, topt in (("config2_hard", False, train.TrainOpt.hard()), ("config3_soft", True, train.TrainOpt.soft())):
        # forward: the decoders (heads included) and the stride-1 pyramid layers with >= 64 channels on tcgen05 ((hi, lo)
        # weights re-packed from the flat parameters every step, one launch); backward: their input gradients
        # (b2f_conv3x3_tc_backward_data), weight gradients (b2f_conv3x3_tc_backward_weights) and the input gradients
        # of the stride-2 pyramid layers (b2f_conv3x3_tc_backward_data_s2) on tcgen05 too; the rest on the FFMA kernels
        net = pwc.PWCNet(pwc.Opt(past_flow=past_flow), device=dev, image_warps=True, tensor_cores=True, train_planar=True)
        res = {}
        for label, c in ((("with_allreduce", cm),) if world > 1 else ()) + (("local", None),):
            res[label], losses = timed(net, topt, c)
        step_ms = res.get("with_allreduce", res["local"])
        out[key] = {"ms_per_step": round(step_ms, 3), "samples_per_s": round(world * B / step_ms * 1e3, 1),
                    "tensor_cores": "decoders (heads included) and stride-1 pyramid layers with >= 64 channels on tcgen05 in "
                                    "forward, input gradient and weight gradient, stride-2 pyramid layers in the input "
                                    "gradient (three-pass TF32 split); the other pyramid work on the FFMA kernels",
                    "h2d_bytes_per_step": B * 9 * H * W * 4,
                    "input_upload": "every step from pinned host memory, started inside the previous step (train_batch prefetch)",
                    "loss": round(losses["err"], 4),
                    "parameters": net.n_params(), "flat_gradient_floats": int(net.flat_params.numel())}
        if world > 1:
            out[key]["step_without_allreduce_ms"] = round(res["local"], 3)
            out[key]["allreduce_exposed_ms"] = round(res["with_allreduce"] - res["local"], 3)
        del net
        torch.cuda.empty_cache()
        if world == 1:
            net = pwc.PWCNet(pwc.Opt(past_flow=past_flow), device=dev, image_warps=True)
            ms, _l = timed(net, topt, None)
            out[key]["ms_per_step_ffma_only"] = round(ms, 3)
            del net
            torch.cuda.empty_cache()
    if cm is not None:
        cm.destroy()
    return out


# ----------------------------------------------------------------------------------------
# criterions (BASELINE configs[2], [3]): informational block, not part of `value`
# ----------------------------------------------------------------------------------------

def time_criterions(torch, lib, dev, B=BATCH, iters=9):
    """Device time of the fused criterion calls of a training step at the reference's training pyramid
    (8 x 320 x 640 ... 20 x 40, RoamingImages / KITTI crops): SURVEY rows a6-a14.  L2 is flushed before every
    call.  Returns microseconds summed over the five levels and GB/s (algorithmic bytes) at the top level."""
    from back2future_b200 import _lib
    P = lambda t: C.c_void_p(t.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)

    def timeit(fn):
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(fn())
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    names = ("OBCC", "OBGCC", "Smoothness(flow,L1)", "Smoothness(occ,Quadratic)", "SecondOrderSmoothness", "ConstVel",
             "OcclusionPrior")
    bpp = (84, 84, 28, 28, 28, 32, 16)
    tot = dict.fromkeys(names, 0.0)
    top = {}
    keep = []
    by_level = []      # the calls (and, through their closures, the tensors) of every level, for the back-to-back timing

    def make_level(k, shs=None):
        # shs: stream handles; call i of the level goes to shs[(k + i) % len(shs)] and every call has its own loss slot
        lv = torch.zeros(8, dtype=torch.float64, device=dev)
        keep.append(lv)
        slot = [0]

        def nxt():
            i = slot[0]
            slot[0] += 1
            return C.c_void_p(lv.data_ptr() + 8 * i), (shs[(k + i) % len(shs)] if shs else None)
        h, w = 320 >> k, 640 >> k
        flow, bflow = torch.randn(B, 2, h, w, device=dev) * 0.2, torch.randn(B, 2, h, w, device=dev) * 0.2
        occ = torch.softmax(torch.randn(B, 2, h, w, device=dev), 1).contiguous()
        w1, w2, tgt = (torch.rand(B, 3, h, w, device=dev) * 4.7 - 2.1 for _ in range(3))
        g2a, g2b, g3a, g3b = torch.empty_like(flow), torch.empty_like(flow), torch.empty_like(w1), torch.empty_like(w1)
        calls = []
        for gt in (0, 1):
            prm = _lib.ObParams(gt, 1, 0.05, 1.0, 0.0 if gt else 1.0, 1.0, 1.0, 20.0 / 2 ** k, gt, 0, 0)
            lp, st_ = nxt()
            calls.append(lambda prm=prm, lp=lp, st_=st_: lib.b2f_ob_criterion(
                C.byref(prm), P(flow), P(bflow), P(occ), P(w1), P(w2), P(tgt), B, 3, h, w, P(g2a), P(g3a), P(g3b), lp, None, st_))
        for order, pen, src in ((1, 1, flow), (1, 0, occ), (2, 1, flow)):
            prm = _lib.SmoothParams(order, pen, 0.05, 20.0, 0, 1)
            lp, st_ = nxt()
            gsm = torch.empty_like(flow)      # own gradient buffer: the calls of a level may run concurrently
            calls.append(lambda prm=prm, src=src, lp=lp, st_=st_, gsm=gsm: lib.b2f_smoothness_criterion(
                C.byref(prm), P(src), P(tgt), B, 2, 3, h, w, P(gsm), lp, None, st_))
        lp, st_ = nxt()
        gcv = torch.empty_like(flow)
        calls.append(lambda lp=lp, st_=st_: lib.b2f_constvel_criterion(P(flow), P(bflow), B, 2, h, w, 1, P(gcv), P(g2b), lp, None, st_))
        lp, st_ = nxt()
        gop = torch.empty_like(flow)
        calls.append(lambda lp=lp, st_=st_: lib.b2f_occprior_criterion(P(occ), B, 2, h, w, 1.0, 0, P(gop), lp, None, st_))
        return h, w, calls

    for k in range(5):
        h, w, calls = make_level(k)
        by_level.append(dict(zip(names, calls)))
        for n, fn, bytes_pp in zip(names, calls, bpp):
            t = timeit(fn)
            tot[n] += t
            if k == 0:
                top[n] = round(bytes_pp * B * h * w / t / 1e3, 1)
    cfg3 = tot["OBCC"] + tot["Smoothness(flow,L1)"] + tot["Smoothness(occ,Quadratic)"] + tot["OcclusionPrior"]
    cfg4 = (tot["OBGCC"] + tot["SecondOrderSmoothness"] + tot["Smoothness(occ,Quadratic)"] + tot["ConstVel"]
            + tot["OcclusionPrior"])
    # the same calls the way a training step issues them (train.lua:416-475): all levels back to back on one stream,
    # losses left on the device (loss_dev), L2 flushed once in front -- launch latency and the loss hand-off of one
    # call overlap the next call's kernel instead of being counted once per call
    def step_of(keys):
        def run():
            for lvl in by_level:
                for n in keys:
                    rc = lvl[n]()
                    if rc:
                        raise RuntimeError("criterion call failed: %d" % rc)
            return 0
        return run
    hard = ("Smoothness(flow,L1)", "OBCC", "Smoothness(occ,Quadratic)", "OcclusionPrior")
    soft = ("SecondOrderSmoothness", "ConstVel", "OBGCC", "Smoothness(occ,Quadratic)", "OcclusionPrior")
    b2b3, b2b4 = timeit(step_of(hard)), timeit(step_of(soft))
    # ... and captured once as a CUDA graph (every entry only enqueues work on the stream it is given when the
    # loss stays on the device): the ~20 launches of a step cost one replay
    gss = [torch.cuda.Stream(device=dev) for _ in range(4)]
    gs = gss[0]
    gshs = [C.c_void_p(x.cuda_stream) for x in gss]
    by_level_eager = by_level
    by_level = [dict(zip(names, make_level(k, gshs)[2])) for k in range(5)]
    graphs = {}
    for key, keys in (("hard", hard), ("soft", soft)):
        run = step_of(keys)
        torch.cuda.synchronize()
        run()                          # first call on every stream: kernel attributes, scratch buffers
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=gs):
            # the criterion calls of a step are independent of each other (each reads the level's outputs and
            # writes its own gradient): four streams, joined at the end
            for x in gss[1:]:
                x.wait_stream(gs)
            run()
            for x in gss[1:]:
                gs.wait_stream(x)
        graphs[key] = gr
    g3 = timeit(lambda: (graphs["hard"].replay(), 0)[1])
    g4 = timeit(lambda: (graphs["soft"].replay(), 0)[1])
    by_level = by_level_eager
    return {"workload": "fused criterion calls (loss + gradients) of one training step, B=8, 320x640 .. 20x40, L2 flushed",
            "us_sum_over_5_levels": {n: round(v, 1) for n, v in tot.items()},
            "GBps_alg_at_320x640": top,
            "config3_hard_step_us": round(cfg3, 1), "config4_soft_step_us": round(cfg4, 1),
            "config3_hard_step_back_to_back_us": round(b2b3, 1), "config4_soft_step_back_to_back_us": round(b2b4, 1),
            "config3_hard_step_graph_us": round(g3, 1), "config4_soft_step_graph_us": round(g4, 1),
            "note": "step_us = sum of the calls timed one by one (L2 flushed before each); back_to_back = the step's calls "
                    "issued in train.lua's order on one stream; graph = the same calls captured once over four streams "
                    "(they are independent of each other) and replayed"}


# ----------------------------------------------------------------------------------------
# e2e: module API, host buffers
# ----------------------------------------------------------------------------------------

def _gpu_local_cpus(torch, dev):
    """CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function), or None."""
    try:
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        bus = None
        pr = torch.cuda.get_device_properties(idx)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        else:
            buf = C.create_string_buffer(32)
            for name in ("libcudart.so.12", "libcudart.so"):
                try:
                    rt = C.CDLL(name)
                    if rt.cudaDeviceGetPCIBusId(buf, 32, C.c_int(idx)) == 0:
                        bus = buf.value.decode().lower()
                        break
                except OSError:
                    continue
        if not bus:
            return None
        if len(bus.split(":")[0]) == 8:      # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        return cpus or None
    except Exception:
        return None


class _NumaLocal:
    """While pinned host buffers are allocated: run on the CPUs next to the GPU, so that the pages are taken from
    that NUMA node (first touch) and the H2D / D2H copies of several ranks do not cross the socket interconnect."""

    def __init__(self, torch, dev):
        self.prev = None
        self.bound = 0
        try:
            cur = os.sched_getaffinity(0)
            local = _gpu_local_cpus(torch, dev)
            want = (local & cur) if local else None
            if want and len(want) < len(cur):
                self.prev = cur
                os.sched_setaffinity(0, want)
                self.bound = len(want)
        except Exception:
            self.prev = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            try:
                os.sched_setaffinity(0, self.prev)
            except Exception:
                pass
        return False


class E2E:
    """The same pass through back2future_b200.nn with pinned HOST inputs and outputs.

    Every step copies every input from pinned host memory and reads every result back.  The tensors of a step live
    back to back in ONE pinned input arena and ONE pinned output arena, so a step is one H2D copy (0.88 GB), the module
    calls, and one D2H copy (0.93 GB) -- 153 separate copies of 7 KB .. 100 MB reached 39 + 41 GB/s, the link does
    46.7 GB/s per direction with both directions busy.  Copies and kernels are pipelined over three streams (H2D /
    compute / D2H) and two sets of device arenas, the way a caller streaming triplets through the reference's modules
    would overlap its data movement: step k + 1 uploads while step k computes and step k - 1 downloads."""

    def __init__(self, torch, dev, B=BATCH, seed=2):
        from back2future_b200 import nn as bnn
        self.torch, self.dev, self.B = torch, dev, B
        g = torch.Generator().manual_seed(seed)
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        # (kind, module(s), input shapes (+ scale), output shapes)
        spec = []
        for l in (7, 6, 5, 4, 3):
            Cn = LEVEL_C[l]
            h, w = level_hw(l)
            spec.append(("cv", (bnn.CostVolMulti(9, True), bnn.CostVolMulti(9, False)),
                         [((B, Cn, h, w), 1.0)] * 3 + [((B, 162, h, w), 1.0)],
                         [(B, 162, h, w)] + [(B, Cn, h, w)] * 4))
        cfgs = [(LEVEL_C[l], level_hw(l)) for l in (6, 5, 4, 3)] + [(3, (H_FULL >> k, W_FULL >> k)) for k in (4, 3, 2, 1, 0)]
        for Cn, (h, w) in cfgs:
            for _ in range(2):
                spec.append(("warp", bnn.BilinearSamplerBHWD(),
                             [((B, h, w, Cn), 1.0), ((B, h, w, 2), 4.0), ((B, h, w, Cn), 1.0)],
                             [(B, h, w, Cn), (B, h, w, Cn), (B, h, w, 2)]))
        def numel(shp):
            n = 1
            for v in shp:
                n *= int(v)
            return n

        n_in = sum(numel(shp) for _k, _m, ins, _o in spec for shp, _sc in ins)
        n_out = sum(numel(shp) for _k, _m, _i, outs in spec for shp in outs)
        self.h2d, self.d2h = 4 * n_in, 4 * n_out
        with _NumaLocal(torch, dev) as numa:
            self.numa_cpus = numa.bound
            self.h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
            self.h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
        self.d_in = [torch.empty(n_in, device=dev, dtype=torch.float32) for _ in range(2)]
        self.d_out = [torch.empty(n_out, device=dev, dtype=torch.float32) for _ in range(2)]
        self.items = []
        oi = oo = 0
        for kind, mod, ins, outs in spec:
            vin = [[], []]
            for shp, sc in ins:
                n = numel(shp)
                self.h_in[oi:oi + n].copy_((torch.randn(shp, generator=g, dtype=torch.float32) * sc).reshape(-1))
                for s_ in range(2):
                    vin[s_].append(self.d_in[s_][oi:oi + n].view(shp))
                oi += n
            vout = [[], []]
            for shp in outs:
                n = numel(shp)
                for s_ in range(2):
                    vout[s_].append(self.d_out[s_][oo:oo + n].view(shp))
                oo += n
            self.items.append((kind, mod, vin, vout))
        self.ev_cmp = [None, None]
        self.ev_out = [None, None]
        self.k = 0

    def step(self):
        """One pass.  Nothing here waits for the host: the three streams are ordered by events only (the input arena of
        a slot is not overwritten before the kernels of the step that used it two steps ago have run; its output arena
        not before that step's read-back has finished).  `drain()` ends the timed region."""
        torch = self.torch
        slot = self.k & 1
        self.k += 1
        with torch.cuda.stream(self.s_in):
            if self.ev_cmp[slot] is not None:
                self.s_in.wait_event(self.ev_cmp[slot])
            self.d_in[slot].copy_(self.h_in, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record()
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(ev_in)
            if self.ev_out[slot] is not None:
                self.s_cmp.wait_event(self.ev_out[slot])
            for kind, mod, vin, vout in self.items:
                din, dout = vin[slot], vout[slot]
                if kind == "cv":
                    ref, past, fut, gj = din
                    joined = dout[0]
                    mod[0].updateOutput([ref, fut], out=joined[:, :81])
                    mod[1].updateOutput([ref, past], out=joined[:, 81:])
                    gf = mod[0].updateGradInput([ref, fut], gj[:, :81])
                    gp = mod[1].updateGradInput([ref, past], gj[:, 81:])
                    res = [None, gf[0], gf[1], gp[0], gp[1]]
                else:
                    img, grid, go = din
                    out = mod.updateOutput([img, grid])
                    gi, gg = mod.updateGradInput([img, grid], go)
                    res = [out, gi, gg]
                for dst, src in zip(dout, res):
                    if src is not None:
                        dst.copy_(src, non_blocking=True)       # module-owned result -> output arena (device to device)
            ev_cmp = torch.cuda.Event()
            ev_cmp.record()
            self.ev_cmp[slot] = ev_cmp
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_cmp)
            self.h_out.copy_(self.d_out[slot], non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record()
            self.ev_out[slot] = ev_out

    def drain(self):
        self.s_out.synchronize()
        self.s_cmp.synchronize()
        self.s_in.synchronize()


# ----------------------------------------------------------------------------------------
# CPU arm: the restatement of the reference algorithm (oracle/c) on host cores
# ----------------------------------------------------------------------------------------

class CpuArm:
    def __init__(self, frac=1.0, seed=2):
        import numpy as np
        path = os.path.join(ROOT, "oracle", "c", "libb2f_cpu.so")
        if not os.path.exists(path):
            import __graft_entry__ as ge
            ge.build_oracle()
        self.lib = C.CDLL(path)
        self.lib.b2fcpu_costvol_backward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                     C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 for N > 1, which would time
        # the reference arm on one core
        try:
            usable = len(os.sched_getaffinity(0))
        except AttributeError:
            usable = os.cpu_count() or 1
        self.lib.b2fcpu_set_num_threads.argtypes = [C.c_int]
        self.lib.b2fcpu_set_num_threads.restype = None
        self.lib.b2fcpu_set_num_threads(usable)
        self.cores = self.lib.b2fcpu_num_threads()
        self.np = np
        rng = np.random.default_rng(seed)
        self.frac = frac
        self.cv, self.warps = [], []
        self.pixels_full = 0
        self.pixels = 0
        f32 = lambda *s: rng.standard_normal(s, dtype=np.float32)
        B = 1
        for l in (7, 6, 5, 4, 3):
            Cn = LEVEL_C[l]
            hf, w = level_hw(l)
            h = max(1, int(round(hf * frac)))
            self.pixels_full += hf * w
            self.pixels += h * w
            ref, past, fut = f32(B, Cn, h, w), f32(B, Cn, h, w), f32(B, Cn, h, w)
            self.cv.append((Cn, h, w, ref, past, fut, np.empty((B, 162, h, w), np.float32), f32(B, 162, h, w),
                            [np.empty((B, Cn, h, w), np.float32) for _ in range(4)]))
        cfgs = [(LEVEL_C[l], level_hw(l)) for l in (6, 5, 4, 3)] + [(3, (H_FULL >> k, W_FULL >> k)) for k in (4, 3, 2, 1, 0)]
        for Cn, (hf, w) in cfgs:
            h = max(1, int(round(hf * frac)))
            for _ in range(2):
                self.warps.append((Cn, h, w, f32(B, h, w, Cn), f32(B, h, w, 2) * 4, f32(B, h, w, Cn),
                                   np.empty((B, h, w, Cn), np.float32), np.empty((B, h, w, Cn), np.float32),
                                   np.empty((B, h, w, 2), np.float32)))

    def step(self):
        L, np = self.lib, self.np
        vp = lambda a: C.c_void_p(a.ctypes.data)
        for Cn, h, w, ref, past, fut, joined, gj, grads in self.cv:
            tmp = np.empty((1, 81, h, w), np.float32)
            for d, (frame, fwd) in enumerate(((fut, 1), (past, 0))):
                ptrs = (C.c_void_p * 2)(ref.ctypes.data, frame.ctypes.data)
                L.b2fcpu_costvol_forward(ptrs, 2, 1, Cn, h, w, 9, fwd, vp(tmp))
                joined[:, 81 * d:81 * (d + 1)] = tmp       # JoinTable copy, as in the reference (pwc.lua:267)
                g = (C.c_void_p * 2)(grads[2 * d].ctypes.data, grads[2 * d + 1].ctypes.data)
                go = gj[:, 81 * d:81 * (d + 1)]
                L.b2fcpu_costvol_backward(ptrs, 2, 1, Cn, h, w, 9, fwd, C.c_void_p(go.ctypes.data), gj.strides[0] // 4, g)
        for Cn, h, w, img, grid, go, out, gi, gg in self.warps:
            L.b2fcpu_warp_forward(vp(img), vp(grid), vp(out), 1, h, w, Cn, h, w)
            gi.fill(0)
            L.b2fcpu_warp_backward(vp(img), vp(grid), vp(go), vp(gi), vp(gg), 1, h, w, Cn, h, w)

    @property
    def triplets_per_step(self):
        return self.pixels / self.pixels_full


def time_cpu(budget_s, steps=None, warmup=1):
    """Time the CPU restatement on a bounded sample.  Returns (triplets/s, cores, description, seconds
    per step, steps)."""
    arm = CpuArm(1.0)
    t0 = time.perf_counter()
    arm.step()                      # calibration / warm-up on one full triplet
    t1 = time.perf_counter() - t0
    n = steps if steps is not None else max(1, min(400, int(budget_s / max(t1, 1e-3))))
    frac = 1.0
    if steps is not None and steps * t1 > budget_s:
        frac = max(0.08, budget_s / (steps * t1))
        arm = CpuArm(frac)
        for _ in range(max(1, warmup)):
            arm.step()
    else:
        for _ in range(max(0, warmup - 1)):
            arm.step()
    t0 = time.perf_counter()
    for _ in range(n):
        arm.step()
    dt = time.perf_counter() - t0
    desc = ("%d step(s) of 1 triplet (B=1) through the full pyramid" % n if frac == 1.0 else
            "%d step(s) of the top %.0f%% of rows of 1 triplet at every pyramid level (%.3f triplet per step)"
            % (n, frac * 100, arm.triplets_per_step))
    return arm.triplets_per_step * n / dt, arm.cores, desc, dt / n, n


# ----------------------------------------------------------------------------------------
# main
# ----------------------------------------------------------------------------------------

def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


def run_reference(args, rank, world):
    if rank != 0:
        return
    value, cores, desc, sps, n = time_cpu(budget_s=120.0, steps=args.steps, warmup=min(args.warmup, 2))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": 1,
                   "batch_note": "the CPU arm runs the same per-triplet work one triplet (B = 1) per step; both arms "
                                 "report triplets/s, so the ratio is per triplet",
                   "note": "Torch7 is not installable here; this is the C/OpenMP restatement of "
                   "the reference algorithm (oracle/c/b2f_cpu.c) on the host cores"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU baseline work (rank 0, N=1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-criterions", action="store_true")
    ap.add_argument("--no-network", action="store_true", help="skip the whole-network (conv trunk included) block")
    ap.add_argument("--no-training", action="store_true", help="skip the training-step block")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed step's outputs")
    ap.add_argument("--breakdown", default=None, help="write the per-kernel table to this JSON file")
    ap.add_argument("--eager", action="store_true", help="time eager C-ABI calls on one stream instead of the CUDA graph")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly ONE line (the JSON record): anything a library prints there (NCCL's version banner at
    # communicator creation, ...) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from back2future_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    W = max(3, args.warmup)
    K = args.steps

    def barrier():
        if world > 1:
            dist.barrier()

    wl = Workload(torch, lib, dev)
    stream = torch.cuda.current_stream()
    wl.bind(stream.cuda_stream)
    for _ in range(W):
        wl.step()
    torch.cuda.synchronize()

    # launches of one step (counted on an eager pass; the graph replays exactly these)
    lib.b2f_launch_count(1)
    wl.step()
    torch.cuda.synchronize()
    launches_per_step = int(lib.b2f_launch_count(0))

    # ---- timed region: exactly K steps ------------------------------------------------
    # default: each step is one replay of the captured CUDA graph (future / past branches on two streams);
    # --eager: the same calls issued one by one on one stream
    dominant = ("costvol_bwd_L3", "costvol_fwd_L3")
    marks = {k: [] for k in dominant}
    graph = None
    if not args.eager:
        gstreams = [torch.cuda.Stream(device=dev) for _ in range(6)]
        graph = wl.capture(gstreams)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        if graph is not None:
            graph.replay()
        else:
            wl.step(marks)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    launches = launches_per_step * K
    ms = e0.elapsed_time(e1)
    ms_local = ms

    # ---- parity gate on the buffers the timed region has just written (rank 0): all 56 ops of the step, every
    # batch item, against the float64 checker at 1e-4 (oracle/ is the checker here, never the thing measured).
    # A kernel left in a measurement mode (b2f_debug_costvol_path 8-10) or a wrong dispatch fails the run.
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import workload_check
        t_par = time.perf_counter()
        parity = workload_check.check_workload(wl)
        parity["seconds"] = round(time.perf_counter() - t_par, 1)
        parity["tensors_compared"], parity["checked"] = parity["checked"], True
        parity["passed"] = not parity["failed"]
        parity["where"] = ("output buffers of the last timed %s" % ("graph replay" if graph is not None else "eager step"))
    if graph is not None:
        # live roofline sample of the dominant kernels: K more eager steps with events around those calls (the
        # graph has no per-kernel events); these steps are outside the timed region
        for _ in range(min(K, 20)):
            wl.step(marks)
        torch.cuda.synchronize()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = world * BATCH * K / (ms / 1e3)

    # ---- roofline of the dominant kernel (live events from the timed region) -----------
    peak, peak_src = load_peaks()
    roof = None
    best = None
    for kind, lst in marks.items():
        if not lst:
            continue
        ts = [a.elapsed_time(b) for a, b, _ in lst]
        avg = sum(ts) / len(ts)
        share = (sum(ts) / len(ts)) * sum(1 for o_ in wl.ops if o_.kind == kind) / (ms_local / K)
        if best is None or sum(ts) > best[0]:
            best = (sum(ts), kind, avg, lst[0][2], share)
    if best:
        _, kind, avg, op, share = best
        achieved = op.bytes / (avg / 1e3) / 1e9
        traffic = load_traffic().get(kind)
        roof = {"bound": "hbm", "kernel": kind, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": op.bytes, "avg_launch_ms": round(avg, 5),
                "share_of_step": round(share, 4),
                "gflops": round(op.flops / (avg / 1e3) / 1e9, 1) if op.flops else None}

    # ---- e2e through the module API with host buffers ----------------------------------
    e2e = None
    if not args.no_e2e:
        ee = E2E(torch, dev)
        for _ in range(3):
            ee.step()
        ee.drain()
        barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()                       # default stream; every e2e stream starts after this point
        for st_ in (ee.s_in, ee.s_cmp, ee.s_out):
            st_.wait_event(a)
        for _ in range(args.e2e_steps):
            ee.step()
        for st_ in (ee.s_in, ee.s_cmp, ee.s_out):   # the end mark waits for the last read-back
            ev_ = torch.cuda.Event()
            ev_.record(st_)
            torch.cuda.current_stream().wait_event(ev_)
        b.record()
        ee.drain()
        torch.cuda.synchronize()
        barrier()
        ems = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = t.item()
        e2e = {"value": round(world * BATCH * args.e2e_steps / (ems / 1e3), 2), "unit": UNIT,
               "h2d_bytes_per_step": ee.h2d, "d2h_bytes_per_step": ee.d2h, "steps": args.e2e_steps,
               "ms_per_step": round(ems / args.e2e_steps, 3),
               "api": "back2future_b200.nn.CostVolMulti / BilinearSamplerBHWD updateOutput+updateGradInput",
               "copies": "one pinned input arena up and one output arena down per step (every input and every result), "
                         "three streams, two sets of device arenas",
               # CPUs this rank was bound to while it allocated its pinned buffers (GPU-local NUMA node), 0 = not bound
               "pinned_alloc_numa_cpus": ee.numa_cpus}
        del ee

    # ---- training path only: the gradient all-reduce next to / under the hot-path step, N > 1 ----
    # The microbenchmark itself has no collective (triplets are sharded); a training step adds exactly one
    # sum-all-reduce of the flattened fp32 gradient (7.19 M floats for Ours-Hard) on a side stream.  Reported: its
    # time alone, and the step time when it runs under the next step's kernels (outside the timed region).
    allreduce = None
    if world > 1:
        # the collective alone, through libb2f_comm.so (the C ABI a LuaJIT host binds); how much of it a training step
        # exposes is measured on a real backward in the `training` block (allreduce_exposed_ms)
        from back2future_b200 import comm as bcomm
        from back2future_b200.dist import NPARAMS_HARD
        cm = bcomm.Communicator.from_env()
        flat = torch.randn(NPARAMS_HARD, device=dev)
        for _ in range(3):
            cm.allreduce_sum(flat)
        torch.cuda.synchronize()
        barrier()
        ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ta.record()
        for _ in range(20):
            cm.allreduce_sum(flat)
        tb.record()
        torch.cuda.synchronize()
        t = torch.tensor([ta.elapsed_time(tb) / 20], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_alone = t.item()
        nbytes = flat.numel() * 4
        allreduce = {"api": "b2f_comm_allreduce_sum_f32 (libb2f_comm.so -> ncclAllReduce)", "floats": flat.numel(),
                     "bytes": nbytes, "alone_ms": round(t_alone, 4),
                     "busbw_GBps": round(2 * (world - 1) / world * nbytes / (t_alone * 1e-3) / 1e9, 1),
                     "overlap": "see training.*.allreduce_exposed_ms"}
        cm.destroy()
        del flat, cm

    # ---- per-kernel breakdown (informational) and CPU baseline (rank 0) ----------------
    rows = breakdown(torch, wl) if rank == 0 else None
    # the same step with other flow statistics (graph replays; outside the timed region, N = 1 only)
    flow_var = None
    if graph is not None and rank == 0 and world == 1 and not args.no_criterions:
        flow_var = {"iid_sigma4_ms": round(ms / K, 4)}
        for kind, key in (("iid0.5", "iid_sigma0.5_ms"), ("smooth", "smooth_ms")):
            wl.set_flow(kind)
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fa.record()
            for _ in range(30):
                graph.replay()
            fb.record()
            torch.cuda.synchronize()
            flow_var[key] = round(fa.elapsed_time(fb) / 30, 4)
        wl.set_flow("iid4", seed=2)
    crit = time_criterions(torch, lib, dev) if (rank == 0 and world == 1 and not args.no_criterions) else None
    infer = time_inference(torch, lib, dev) if (rank == 0 and world == 1 and not args.no_criterions) else None
    network = None
    if not args.no_network:
        network = time_network(torch, dev, world, rank, dist, cpu=not args.no_cpu)
    training = None
    if not args.no_training:
        training = time_training(torch, dev, world, rank, dist)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, desc, sps, n = time_cpu(args.cpu_budget)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
               "s_per_triplet": round(sps, 3)}

    if rank == 0:
        if rows:
            tot = sum(r["ms"] for r in rows)
            sys.stderr.write("per-kernel (separate pass, median of 10):\n")
            for r in sorted(rows, key=lambda r: -r["ms"])[:24]:
                sys.stderr.write("  %-34s %8.4f ms  %6.1f%%  %8s GB/s  %s\n" % (
                    r["kernel"], r["ms"], 100 * r["ms"] / tot, r["GBps"], ("%s GFLOP/s" % r["GFLOPs"]) if r["GFLOPs"] else ""))
            sys.stderr.write("  sum of kernels %.4f ms vs %.4f ms per step\n" % (tot, ms / K))
            if args.breakdown:
                json.dump(rows, open(args.breakdown, "w"), indent=1)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "sharding": "triplets across ranks, no collective",
                       "l2": "no explicit flush: one step touches %.2f GB of distinct buffers (> 126 MB L2) before any "
                             "buffer is reused" % (wl.total_bytes() / 1e9),
                       "alg_bytes_per_step": wl.total_bytes(),
                       "launch": "eager, one stream" if args.eager else
                                 "one CUDA-graph replay per step; the graph's edges are the network's data dependencies "
                                 "(pwc.lua:237-456): per level [CV future || CV past] -> [feature warp future || past] -> next "
                                 "level, the level's image warps (leaves of the loss) on side streams under the next level's "
                                 "cost volumes, backward mirrored; the gradImg zero-fills (b2f_zero_async) are issued at the "
                                 "start of the step on their own stream"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "parity": parity,
            # every hot kernel against the same HBM peak (separate pass with events around each call, one call at a
            # time: the isolated times behind the stderr table; `roofline` above is the cost-volume kernel BASELINE's
            # metric names, the C = 3 image-warp backward is as long per launch)
            "roofline_by_kernel": ([{"kernel": r["kernel"], "ms": r["ms"], "GBps": r["GBps"],
                                     "frac": round(r["GBps"] / roof["peak"], 4) if (r["GBps"] and roof and roof.get("peak")) else None}
                                    for r in sorted(rows, key=lambda r: -r["ms"])[:12]] if rows else None),
            "cpu_baseline": cpu,
            "criterions": crit, "flow_variants": flow_var, "inference_shapes": infer, "allreduce": allreduce,
            "whole_network": network, "training": training,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and parity["failed"]:
        sys.stderr.write("PARITY FAILED: %r\n" % (parity,))
        sys.exit(3)


if __name__ == "__main__":
    main()
