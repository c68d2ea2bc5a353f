"""`trainBatch` of the reference (train.lua:196-520, the `optimize == 'pme'` unsupervised branch) over the B200 path:
forward, the occlusion-aware photometric / smoothness / prior criterions at every output level, `model:backward`,
the gradient all-reduce of the data-parallel run and `optim.adam`.

One TRAINING STEP is a fixed list of C-ABI calls over preallocated buffers (captured once as a CUDA graph):

    inputs (B, 9, H, W) -> PWCNet forward plan (pwc.py)
    reference-frame pyramid  down_sampled[{{}, {4, 6}}]  by repeated SpatialAveragePooling(2,2,2,2)   train.lua:281-283, 419
    per level l = 0 (finest) .. 4, weights level_weights[l+1] = {0.005, 0.01, 0.02, 0.08, 0.32}        train.lua:56-58
        flow smoothness      lw * smooth_flow * fs_criterion(flow[, bflow], target)                   :427-433
        constant velocity    lw * const_vel * cv_criterion                       (past_flow models)   :436-441
        photometric          lw * pme * pme_criterion({flow,[bflow,]occ,w1,w3}, target)               :444-454
        occlusion smoothness lw * smooth_occ * os_criterion(occ, target)                              :458-462
        occlusion prior      lw * prior_occ * oprior_criterion(occ)                                   :465-468
      each criterion is ONE fused kernel giving loss + gradient; the weighted gradients are accumulated into the
      gradOutputs table with b2f_axpy2d, the losses stay on the device (one double each) until the step ends
    PWCNet backward plan, all-reduce(sum) of the flat gradient (libb2f_comm.so), Adam on the flat parameters.

Data-parallel semantics: every rank runs the criterions on its own shard; `sizeAverage = false` losses are
sums, so gradients add across ranks with no rescale.  The Q9 weight aliasing of the first-order smoothness criterion
couples samples of one criterion call, i.e. of one rank's shard.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, pwc

LEVEL_WEIGHTS = (0.005, 0.01, 0.02, 0.08, 0.32, 0.64, 1.28)          # train.lua:56-58
_PEN = {"Quadratic": (_lib.PENALTY_QUADRATIC, 0.0), "L1": (_lib.PENALTY_L1, 0.0),
        "Lorentzian": (_lib.PENALTY_LORENTZIAN, 0.05), "Dirac": (_lib.PENALTY_LORENTZIAN, 0.001)}


class TrainOpt:
    """The opts.lua fields trainBatch / model.lua read, reference defaults (opts.lua:57-79)."""

    def __init__(self, **kw):
        self.sizeAverage = False
        self.pme = 1.0
        self.pme_criterion = "OBCC"
        self.pme_penalty = "L1"
        self.pme_alpha = 1.0
        self.pme_beta = 1.0
        self.smooth_flow = 1.0
        self.smooth_second_order = False
        self.smooth_flow_penalty = "L1"
        self.smooth_occ_penalty = "Quadratic"
        self.smooth_occ = 0.1
        self.prior_occ = 0.1
        self.const_vel = 1.0
        self.LR = 1e-4
        self.weightDecay = 0.0
        self.beta1, self.beta2, self.epsilon = 0.9, 0.999, 1e-8      # optim.adam defaults
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown training option %r" % k)
            setattr(self, k, v)
        if self.pme_criterion not in ("OBCC", "OBGCC"):
            raise NotImplementedError("pme_criterion %r: only OBCC and OBGCC are on the hot path" % self.pme_criterion)

    @classmethod
    def hard(cls, **kw):
        """README.md:85-86: hard-constraint pretraining (pme 1, OBCC, smooth_flow 2)."""
        return cls(**dict(dict(pme=1.0, pme_criterion="OBCC", smooth_flow=2.0), **kw))

    @classmethod
    def soft(cls, **kw):
        """README.md:91-94: soft-constraint fine-tuning (pme 2, OBGCC alpha 0, second-order smoothness, const_vel)."""
        return cls(**dict(dict(pme=2.0, pme_criterion="OBGCC", pme_alpha=0.0, pme_beta=1.0, smooth_flow=0.1,
                               smooth_second_order=True, const_vel=1e-4), **kw))


def _vp(t):
    return C.c_void_p(t.data_ptr())


class Trainer:
    LOSSES = ("sflow", "cvel", "pme", "socc", "gocc")

    def __init__(self, net: pwc.PWCNet, opt: TrainOpt | None = None, comm=None):
        if not net.image_warps or (net.tensor_cores and not net.train_planar):
            raise ValueError("training needs the warped frames and the planar activations: build the model with "
                             "image_warps=True and tensor_cores=False (or tensor_cores=True, train_planar=True)")
        self.net, self.opt, self.comm = net, opt or TrainOpt(), comm
        self.lib = net.lib
        self._steps = {}
        self._comm_stream = torch.cuda.Stream(net.device) if comm is not None and comm.world > 1 else None
        self._copy_stream = None
        self._staged, self._stage_buf = {}, {}
        self.batchNumber = 0

    # ---- the step's criterion calls -----------------------------------------------------------------------
    def _build(self, B, H, W):
        net, lib, o = self.net, self.lib, self.opt
        dev = net.device
        p = net.plan(B, H, W)
        if p.bops is None:
            net._build_backward(p)
        mo = net.opt
        nlev = mo.levels - mo.l_st + 1
        per = net.n_unit_out
        nflow = 2 if net.past_flow else 1
        E = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        st = _Step()
        st.plan = p
        ops = st.ops
        # reference frame at every level: channels 4..6 of the progressively average-pooled input
        tg = [E(B, 3, H >> k, W >> k) for k in range(nlev)]
        ops.append((lib.b2f_copy2d_async, (_vp(tg[0]), 3 * H * W, C.c_void_p(p.x.data_ptr() + 4 * 3 * H * W), 9 * H * W,
                                           3 * H * W, B)))
        for k in range(1, nlev):
            ops.append((lib.b2f_avgpool2x2_forward, (_vp(tg[k - 1]), _vp(tg[k]), B, 3, H >> (k - 1), W >> (k - 1))))
        st.targets = tg
        st.loss_dev = torch.zeros(nlev * len(self.LOSSES) * 2, device=dev, dtype=torch.float64)
        st.loss_w = [0.0] * (nlev * len(self.LOSSES) * 2)
        slot = [0]

        def loss_slot(weight):
            i = slot[0]
            slot[0] += 1
            st.loss_w[i] = weight
            return C.c_void_p(st.loss_dev.data_ptr() + 8 * i), i

        for g in p.gout:
            ops.append((lib.b2f_zero_async, (_vp(g), g.numel() * 4)))
        sa = int(bool(o.sizeAverage))
        fs_kind, fs_eps = _PEN[o.smooth_flow_penalty]
        os_kind, os_eps = _PEN[o.smooth_occ_penalty]
        pm_kind, pm_eps = _PEN[o.pme_penalty]
        st.keep = []
        st.names = []
        for k in range(nlev):
            lw = 1.0 if o.sizeAverage else LEVEL_WEIGHTS[k]
            h, w = H >> k, W >> k
            unit = p.output[k * per:(k + 1) * per]
            gunit = p.gout[k * per:(k + 1) * per]
            occ, gocc = unit[nflow], gunit[nflow]
            n2 = B * 2 * h * w

            def add(dst, src, alpha, n):
                ops.append((lib.b2f_axpy2d, (_vp(dst), n, _vp(src), n, n, 1, C.c_float(alpha))))

            # flow smoothness (train.lua:427-433)
            prm = _lib.SmoothParams(2 if o.smooth_second_order else 1, fs_kind, fs_eps, 20.0, sa, 1)
            st.keep.append(prm)
            for i in range(nflow):
                g = E(B, 2, h, w)
                ld, si = loss_slot(lw * o.smooth_flow)
                st.names.append(("sflow", si))
                ops.append((lib.b2f_smoothness_criterion, (C.byref(prm), _vp(unit[i]), _vp(tg[k]), B, 2, 3, h, w, _vp(g), ld,
                                                           None)))
                add(gunit[i], g, lw * o.smooth_flow, n2)
                st.keep.append(g)
            # constant velocity (:436-441)
            if net.past_flow:
                g0, g1 = E(B, 2, h, w), E(B, 2, h, w)
                ld, si = loss_slot(lw * o.const_vel)
                st.names.append(("cvel", si))
                ops.append((lib.b2f_constvel_criterion, (_vp(unit[0]), _vp(unit[1]), B, 2, h, w, sa, _vp(g0), _vp(g1), ld, None)))
                add(gunit[0], g0, lw * o.const_vel, n2)
                add(gunit[1], g1, lw * o.const_vel, n2)
                st.keep += [g0, g1]
            # photometric (:444-454); pwc_flow_scaling = model.flow_scale[levels - l] (:425)
            scaling = net.flow_scale[nlev - 1 - k]
            ob = _lib.ObParams(1 if o.pme_criterion == "OBGCC" else 0, pm_kind, pm_eps, 1.0, float(o.pme_alpha),
                               float(o.pme_beta), 1.0, float(scaling), int(net.past_flow), 0, sa)
            st.keep.append(ob)
            go, gw1, gw3 = E(B, 2, h, w), E(B, 3, h, w), E(B, 3, h, w)
            ld, si = loss_slot(lw * o.pme)
            st.names.append(("pme", si))
            ops.append((lib.b2f_ob_criterion, (C.byref(ob), _vp(unit[0]), _vp(unit[1]) if net.past_flow else None, _vp(occ),
                                               _vp(unit[nflow + 1]), _vp(unit[nflow + 2]), _vp(tg[k]), B, 3, h, w,
                                               _vp(go), _vp(gw1), _vp(gw3), ld, None)))
            add(gocc, go, lw * o.pme, n2)
            add(gunit[nflow + 1], gw1, lw * o.pme, B * 3 * h * w)
            add(gunit[nflow + 2], gw3, lw * o.pme, B * 3 * h * w)
            st.keep += [go, gw1, gw3]
            # occlusion smoothness (:458-462) and prior (:465-468)
            if o.smooth_occ > 0:
                prm2 = _lib.SmoothParams(1, os_kind, os_eps, 20.0, sa, 1)
                g = E(B, 2, h, w)
                ld, si = loss_slot(lw * o.smooth_occ)
                st.names.append(("socc", si))
                ops.append((lib.b2f_smoothness_criterion, (C.byref(prm2), _vp(occ), _vp(tg[k]), B, 2, 3, h, w, _vp(g), ld, None)))
                add(gocc, g, lw * o.smooth_occ, n2)
                st.keep += [prm2, g]
            if o.prior_occ > 0:
                g = E(B, 2, h, w)
                ld, si = loss_slot(lw * o.prior_occ)
                st.names.append(("gocc", si))
                ops.append((lib.b2f_occprior_criterion, (_vp(occ), B, 2, h, w, C.c_float(1.0), sa, _vp(g), ld, None)))
                add(gocc, g, lw * o.prior_occ, n2)
                st.keep.append(g)
        st.loss_host = torch.zeros(st.loss_dev.numel(), dtype=torch.float64).pin_memory()
        return st

    def _segments(self, st):
        """The step as a list of (launch function, gradient bucket or None): with more than one rank the backward plan
        is cut after every point where a range of the flat gradient becomes final, so that its all-reduce runs on the
        communication stream under the rest of the backward (train.lua's DataParallelTable reduces after the whole
        backward).  One rank: a single segment."""
        p = st.plan

        def head():
            p.launch()
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for fn, args in st.ops:
                _lib.check(fn(*args, s))

        if self.comm is None or self.comm.world == 1:
            return [(lambda: (head(), p.launch_backward()), None)]
        segs, lo = [], 0
        nb = len(p.bucket_marks)
        for k, (idx, rng) in enumerate(p.bucket_marks):
            a, b = lo, idx
            last = k == nb - 1       # only the last slice joins the plan's side streams into the main one
            if k == 0:
                segs.append((lambda a=a, b=b, last=last: (head(), p.launch_backward(a, b, join=last)), rng))
            else:
                segs.append((lambda a=a, b=b, last=last: p.launch_backward(a, b, join=last), rng))
            lo = idx
        assert lo == len(p.bops)
        return segs

    def _launch(self, st):
        """forward plan -> criterions -> backward plan, on the current stream (+ the forward plan's side lanes)."""
        for fn, _rng in self._segments(st):
            fn()

    # ---- trainBatch ---------------------------------------------------------------------------------------
    def train_batch(self, inputs, graph=True, step=True, prefetch=None):
        """One optimisation step on `inputs` (B, 9, H, W) (host or device float32).  Returns the weighted losses of
        THIS rank's shard {err, pme, sflow, socc, gocc, cvel} (train.lua:471, 497-517).  `step=False` stops after the
        gradient (flat_grads), for tests.

        `prefetch`: the NEXT step's inputs (a pinned host tensor of the same shape, not modified until that step has
        been issued).  Their upload is started on a copy stream as soon as this step's work is enqueued, into a staging
        buffer, and the next `train_batch` called with that very tensor takes the staged copy (one device-to-device
        copy) instead of uploading again: the 59 MB upload of a step (1.1 ms over PCIe, a twelfth of the step) runs under
        the previous step's kernels -- the double buffering that the reference's loader threads give `trainBatch`
        (train.lua:168-189, donkeys) on the host side."""
        net = self.net
        B, nine, H, W = inputs.shape
        key = (B, H, W)
        with torch.cuda.device(net.device):
            if key not in self._steps:
                self._steps[key] = self._build(B, H, W)
                # the library's loss scratch must exist before a capture
                _lib.check(self.lib.b2f_reserve_scratch(1 << 20))
            st = self._steps[key]
            cur = torch.cuda.current_stream()
            pf = self._staged.pop(key, None)
            if pf is not None and pf[2] is inputs:
                cur.wait_event(pf[1])
                st.plan.x.copy_(pf[0], non_blocking=True)
            else:
                st.plan.x.copy_(inputs, non_blocking=True)
            x_taken = torch.cuda.Event()
            x_taken.record(cur)
            segs = self._segments(st)
            multi = self.comm is not None and self.comm.world > 1
            if graph and st.graph is None:
                self._launch(st)                                       # warm-up: attributes, tensor maps, scratch
                if multi:
                    # NCCL allocates its channels on a communicator's first collective: not inside a capture
                    self.comm.allreduce_sum(net.flat_grads, 0, min(1024, net.flat_grads.numel()), stream=cur)
                cur.synchronize()
                # ONE graph for the whole step.  With more than one rank the bucket all-reduces are captured INTO it
                # (NCCL records its kernels on a capturing stream): the communication stream forks from the capture
                # stream behind the kernels that make a bucket final and joins before the end, so the reductions
                # are graph nodes that run under the rest of the backward -- no host round trip per bucket (ten graph
                # launches and ten event waits per step before: 0.94 ms exposed at 8 GPUs for a 0.14 ms collective).
                g = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream(net.device)
                cap.wait_stream(cur)
                with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                    for fn, rng in segs:
                        fn()
                        if multi and rng is not None:
                            # the bucket is final once the main stream AND the plan's side streams (weight gradients,
                            # flow-decoder chains) have reached this point; only the communication stream waits
                            for s_ in [cap] + st.plan.backward_streams():
                                ev = torch.cuda.Event()
                                ev.record(s_)
                                self._comm_stream.wait_event(ev)
                            self.comm.allreduce_sum(net.flat_grads, rng[0], rng[1], stream=self._comm_stream)
                    if multi:
                        cap.wait_stream(self._comm_stream)
                st.graph = g
            if graph:
                st.graph.replay()
            else:
                for fn, rng in segs:
                    fn()
                    if multi and rng is not None:
                        for s_ in [cur] + st.plan.backward_streams():
                            ev = torch.cuda.Event()
                            ev.record(s_)
                            self._comm_stream.wait_event(ev)
                        self.comm.allreduce_sum(net.flat_grads, rng[0], rng[1], stream=self._comm_stream)
                if multi:
                    cur.wait_stream(self._comm_stream)
            if step:
                net.adam_step(self.opt.LR, self.opt.beta1, self.opt.beta2, self.opt.epsilon, self.opt.weightDecay)
            if prefetch is not None:
                if tuple(prefetch.shape) != tuple(inputs.shape) or prefetch.is_cuda:
                    raise ValueError("train_batch: prefetch must be a host tensor shaped like inputs")
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(net.device)
                buf = self._stage_buf.get(key)
                if buf is None:
                    buf = self._stage_buf[key] = torch.empty(inputs.shape, device=net.device, dtype=torch.float32)
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(x_taken)        # the staging buffer has been read by this step
                    buf.copy_(prefetch, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                self._staged[key] = (buf, ev, prefetch)
            st.loss_host.copy_(st.loss_dev, non_blocking=True)
            cur.synchronize()                                          # cutorch.synchronize(), train.lua:498
        self.batchNumber += 1
        vals = st.loss_host.tolist()
        out = {k: 0.0 for k in self.LOSSES}
        for name, i in st.names:
            out[name] += st.loss_w[i] * vals[i]
        out["err"] = sum(out[k] for k in self.LOSSES)
        return out


class _Step:
    def __init__(self):
        self.ops = []
        self.graph = None
