"""The multi-frame PWC network of the reference as a graph executor over libb2f_cuda.so.

Mirrors `createModelMulti(opt)` of models/pwc.lua:87-508 for the Ours-Hard / Ours-Soft family (frames 3,
two_frame 0, pwc_sum_cvs false, residual 0, occ_input 0, rescale_flow 0, siamese 1, skip > 0): same inputs
((B, 9, H, W), three ColorNormalized RGB frames stacked on the channel axis), same output table (finest level first:
{flow, [past flow,] occlusion, warped frame 1, warped frame 3} per level, pwc.lua:459-489) and the same
`flow_scale` / `past_flow` fields (:493-494).  nngraph's node-by-node interpreter is replaced by a PLAN: the list of
C-ABI calls of one forward pass, built once per input shape over preallocated buffers and replayed either call by
call or as one CUDA graph.  What disappears relative to the reference's graph:

  * nn.JoinTable(2) (:267, 298-305, 334): both cost volumes, the reference features (second destination of the
    convolution that produces them) and the up-sampled flow are written straight into the decoders' joined input;
  * nn.Narrow + the siamese clones (:141-146, 186-195): the three frames run through the shared convUnits as one
    batch of 3B (slot order past, future, reference);
  * nn.Transpose x 3 + nn.MulConstant around every sampler (:68-73, 404, 443): b2f_warp_bdhw_forward;
  * the second decoder input of the Soft models ({cvs, ref, ubfs}, :336): the joined buffer carries BOTH up-sampled
    flows and each decoder's first convolution has zero weights for the one it does not see.

torch owns device memory, streams, events and the CUDA graph object -- nothing else; every kernel is ours.  There
is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib

FEAT = (3, 16, 32, 64, 96, 128, 192)      # featMaps (pwc.lua:89, d = 16)
DEC = (128, 128, 96, 64, 32, 2)           # decoder(nChannels) widths (pwc.lua:76-85)
SIDE_LANES = True                         # backward plan on several streams (tools/time_train.py --no-side measures without)
TC_FEAT_MIN = 64                          # feature-pyramid layers with at least this many channels run on tcgen05


class Opt:
    """The opts.lua fields createModelMulti reads (pwc.lua:100-113), reference defaults (opts.lua:83-98)."""

    def __init__(self, **kw):
        self.pwc_ws = 9
        self.frames = 3
        self.levels = 7
        self.pwc_skip = 2
        self.flownet_factor = 20
        self.past_flow = False
        self.two_frame = 0
        self.pwc_sum_cvs = False
        self.residual = 0
        self.occ_input = 0
        self.rescale_flow = 0
        self.pwc_siamese = 1
        self.nGPU = 1
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("createModelMulti: unknown option %r" % k)
            setattr(self, k, v)
        if (self.frames != 3 or self.two_frame or self.pwc_sum_cvs or self.residual or self.occ_input
                or self.rescale_flow or self.pwc_siamese != 1 or self.pwc_skip < 1):
            raise NotImplementedError("only the published model family is built (frames 3, two_frame 0, pwc_sum_cvs "
                                      "false, residual 0, occ_input 0, rescale_flow 0, pwc_siamese 1, pwc_skip >= 1)")

    @property
    def l_st(self):
        return max(self.pwc_skip + 1, 1)


def conv_shapes(opt):
    """Ordered (name, Cout, Cin, stride) of every convolution: feat.l<l>.<0|1> (shared convUnits, pwc.lua:176-195),
    occ / flow / bflow .l<l>.<0..5> (decoders, :288-338)."""
    shapes = []
    for l in range(2, opt.levels + 1):
        shapes.append(("feat.l%d.0" % l, FEAT[l - 1], FEAT[l - 2], 2))
        shapes.append(("feat.l%d.1" % l, FEAT[l - 1], FEAT[l - 1], 1))
    nd = 2 * opt.pwc_ws ** 2
    for l in range(opt.levels, opt.l_st - 1, -1):
        n_occ = nd + FEAT[l - 1] + (2 if l != opt.levels else 0)
        n_flow = nd if l == opt.levels else nd + FEAT[l - 1] + 2
        kinds = (("occ", n_occ), ("flow", n_flow)) + ((("bflow", n_flow),) if opt.past_flow else ())
        for kind, n_in in kinds:
            cin = n_in
            for i, cout in enumerate(DEC):
                shapes.append(("%s.l%d.%d" % (kind, l, i), cout, cin, 1))
                cin = cout
    return shapes


def _vp(t):
    return C.c_void_p(t.data_ptr())


class _Conv:
    __slots__ = ("w", "b", "cin", "cout", "stride", "gw", "gb", "wt", "w_off", "b_off", "tc_h", "tc_l", "tc_cin", "tct_h", "tct_l")


def _round64(n):
    return (n + 63) // 64 * 64


def _round32(n):
    return (n + 31) // 32 * 32


class PWCNet:
    """`createModelMulti(opt)` + `:cuda()`.  `params`: dict name -> array in Torch's layout ((Cout, Cin, 3, 3) weight,
    (Cout,) bias) with the names of `conv_shapes`; None = nn.SpatialConvolution:reset()'s uniform(-1/sqrt(9 nIn), ..)
    drawn from numpy's default_rng(seed)."""

    def __init__(self, opt=None, params=None, device="cuda:0", seed=2, image_warps=True, tensor_cores=False,
                 train_planar=False):
        """tensor_cores=True: the decoders (five wide convolutions and the 2-channel head) and the stride-1 pyramid layers
        with >= TC_FEAT_MIN channels run on tcgen05 (b2f_conv3x3_tc_forward, three-pass TF32 split, fp32-level accuracy)
        over channel-minor (hi, lo) activations.  train_planar=True is the training configuration of that path (the
        name is from the time it also kept planar copies): the forward plan starts by re-packing the (hi, lo) weights
        from the flat parameter buffer (Adam updates it every step) and records every layer's (hi, lo) input
        (plan.dec_hl), which is all the tensor-core backward plan reads."""
        if not torch.cuda.is_available():
            raise RuntimeError("PWCNet: no CUDA device; the B200 path has no CPU fallback")
        self.opt = opt or Opt()
        self.device = torch.device(device)
        self.lib = _lib.load()
        self.image_warps = bool(image_warps)
        self.tensor_cores = bool(tensor_cores)
        self.train_planar = bool(train_planar) and self.tensor_cores
        o = self.opt
        self.past_flow = bool(o.past_flow)                                             # model.past_flow, pwc.lua:494
        self.flow_scale = [o.flownet_factor / 2.0 ** (l - o.l_st) for l in range(o.levels, o.l_st - 1, -1)]   # :451-455
        self.n_unit_out = 5 if self.past_flow else 4
        self._plans = {}
        self._convs = {}
        if params is None:
            params = self.random_params(self.opt, seed)
        self.load_params(params)

    # ---- parameters -------------------------------------------------------------------------------------
    @staticmethod
    def random_params(opt, seed=2, scale=1.0):
        rng = np.random.default_rng(seed)
        params = {}
        for name, cout, cin, _s in conv_shapes(opt):
            s = scale / math.sqrt(9.0 * cin)
            params[name + ".weight"] = rng.uniform(-s, s, (cout, cin, 3, 3)).astype(np.float32)
            params[name + ".bias"] = rng.uniform(-s, s, (cout,)).astype(np.float32)
        return params

    def n_params(self):
        return sum(co * ci * 9 + co for _n, co, ci, _s in conv_shapes(self.opt))

    def load_params(self, params):
        """Upload and repack every convolution ((Cout, Cin, 3, 3) -> [Cin * 9][CoutP]).  For the Soft models the first
        convolution of each level < levels decoder is widened to the shared joined input {cvs, ref, ufs, ubfs} with zero
        weights on the up-sampled flow it does not read."""
        o = self.opt
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            # ONE flat buffer for every packed weight and bias (64-float aligned pieces): what getParameters() is to
            # the reference (train.lua:24) -- the gradient, the Adam moments and the all-reduce use the same layout
            sizes = []
            for name, cout, cin, stride in conv_shapes(o):
                kind, lvl, idx = name.split(".")
                wide = self.past_flow and idx == "0" and kind in ("occ", "flow", "bflow") and int(lvl[1:]) != o.levels
                sizes.append((int(self.lib.b2f_conv3x3_packed_floats(cin + (2 if wide else 0), cout)), _round64(cout)))
            total = sum(a + b for a, b in sizes)
            if getattr(self, "flat_params", None) is None or self.flat_params.numel() != total:
                self.flat_params = torch.empty(total, device=self.device, dtype=torch.float32)
                _lib.check(self.lib.b2f_zero_async(_vp(self.flat_params), total * 4, st))
            off = 0
            for (name, cout, cin, stride), (nw, nb) in zip(conv_shapes(o), sizes):
                w = np.ascontiguousarray(np.asarray(params[name + ".weight"], np.float32))
                b = np.ascontiguousarray(np.asarray(params[name + ".bias"], np.float32))
                if w.shape != (cout, cin, 3, 3) or b.shape != (cout,):
                    raise ValueError("%s: expected weight %r / bias %r, got %r / %r" %
                                     (name, (cout, cin, 3, 3), (cout,), w.shape, b.shape))
                kind, lvl, idx = name.split(".")
                l = int(lvl[1:])
                if self.past_flow and idx == "0" and kind in ("occ", "flow", "bflow") and l != o.levels:
                    wide = np.zeros((cout, cin + 2, 3, 3), np.float32)
                    wide[:, :cin - 2] = w[:, :cin - 2]
                    if kind == "bflow":
                        wide[:, cin:cin + 2] = w[:, cin - 2:]
                    else:
                        wide[:, cin - 2:cin] = w[:, cin - 2:]
                    w, cin = wide, cin + 2
                cv = _Conv()
                cv.cin, cv.cout, cv.stride = cin, cout, stride
                cv.w_off, cv.b_off = off, off + nw
                wt = torch.from_numpy(w).to(self.device)
                cv.w = self.flat_params[off:off + nw]
                assert nw == int(self.lib.b2f_conv3x3_packed_floats(cin, cout))
                _lib.check(self.lib.b2f_conv3x3_pack_weights(_vp(wt), _vp(cv.w), cout, cin, 0, st))
                cv.b = self.flat_params[off + nw:off + nw + cout]
                cv.b.copy_(torch.from_numpy(b), non_blocking=False)        # cudaMemcpy H2D
                cv.gw = cv.gb = cv.wt = None
                cv.tc_h = cv.tc_l = cv.tct_h = cv.tct_l = None
                if self.tensor_cores and (kind in ("occ", "flow", "bflow") or
                                          (kind == "feat" and idx == "1" and cout >= TC_FEAT_MIN)):
                    # [9][Cout][Cin_p] hi / lo for the tensor-core path.  The coarsest level's flow decoder reads the
                    # first 162 channels of the occlusion decoder's wider joined input: zero weights for the rest.
                    wtc = w
                    if idx == "0" and l == o.levels and kind != "occ":
                        wtc = np.zeros((cout, 2 * o.pwc_ws ** 2 + FEAT[l - 1], 3, 3), np.float32)
                        wtc[:, :cin] = w
                    cv.tc_cin = wtc.shape[1]
                    n_tc = int(self.lib.b2f_conv3x3_tc_packed_floats(cv.tc_cin, cout))
                    cv.tc_h = torch.empty(n_tc, device=self.device, dtype=torch.float32)
                    cv.tc_l = torch.empty(n_tc, device=self.device, dtype=torch.float32)
                    wsrc = torch.from_numpy(np.ascontiguousarray(wtc)).to(self.device)
                    _lib.check(self.lib.b2f_conv3x3_tc_pack_weights(_vp(wsrc), _vp(cv.tc_h), _vp(cv.tc_l), cout,
                                                                     cv.tc_cin, st))
                    torch.cuda.current_stream(self.device).synchronize()
                off += nw + nb
                self._convs[name] = cv
            torch.cuda.current_stream(self.device).synchronize()

    def state_params(self):
        """Every convolution back in Torch's layout (dict of numpy arrays; the Soft models' widened first decoder
        convolutions are narrowed again): what `torch.save(model)` would hold."""
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        out = {}
        with torch.cuda.device(self.device):
            for name, cout, cin0, _s in conv_shapes(self.opt):
                cv = self._convs[name]
                w = torch.empty((cout, cv.cin, 3, 3), device=self.device)
                _lib.check(self.lib.b2f_conv3x3_pack_weights(_vp(w), _vp(cv.w), cout, cv.cin, 1, st))
                w = w.cpu().numpy()
                if cv.cin != cin0:
                    kind = name.split(".")[0]
                    w = np.concatenate([w[:, :cin0 - 2], w[:, cin0:cin0 + 2] if kind == "bflow" else w[:, cin0 - 2:cin0]], 1)
                out[name + ".weight"] = np.ascontiguousarray(w)
                out[name + ".bias"] = cv.b.cpu().numpy().copy()
        return out

    # ---- plan -------------------------------------------------------------------------------------------
    def _build(self, B, H, W):
        o, lib, dev = self.opt, self.lib, self.device
        levels, l_st, win = o.levels, o.l_st, o.pwc_ws
        if H % (1 << (levels - 1)) or W % (1 << (levels - 1)):
            raise ValueError("PWCNet: input %dx%d is not a multiple of %d (back2future.lua:54-67 rescales to one)" %
                             (W, H, 1 << (levels - 1)))
        nd = win * win
        E = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        plan = _Plan(B, H, W)
        x = plan.x = E(B, 9, H, W)
        hw = lambda l: (H >> (l - 1), W >> (l - 1))
        ops = plan.ops
        P = _vp

        def conv(name, xin, xbs, xb, cin, h, w, out, obs, out2=None, obs2=0, slope=0.2, lane=0):
            cv = self._convs[name]
            assert cv.cin == cin, (name, cv.cin, cin)
            ops.append((lane, lib.b2f_conv3x3_forward,
                        (xin, xbs, P(cv.w), P(cv.b), out, obs, out2, obs2, xb, cin, h, w, cv.cout, cv.stride,
                         C.c_float(slope))))

        def sl(t, b0=0, c0=0):
            """device pointer of t[b0, c0]"""
            return C.c_void_p(t.data_ptr() + 4 * (b0 * t.stride(0) + c0 * t.stride(1)))

        # -- joined decoder inputs J[l]: {cvs_fwd 81, cvs_bwd 81, ref features C_l, [ufs 2, [ubfs 2]]} ------------
        nfl = 4 if self.past_flow else 2
        J = {}
        for l in range(l_st, levels + 1):
            h, w = hw(l)
            J[l] = E(B, 2 * nd + FEAT[l - 1] + (nfl if l != levels else 0), h, w)
        plan.J = J

        # -- frames 1 and 3 as dense images + their average-pooled pyramid (ds, pwc.lua:149-158) ------------------
        n_ds = levels - l_st + 1
        ds = [E(2 * B, 3, H >> k, W >> k) for k in range(n_ds)] if self.image_warps else []
        if self.image_warps:
            for i, f in enumerate((0, 2)):
                ops.append((1, lib.b2f_copy2d_async, (sl(ds[0], i * B), 3 * H * W, sl(x, 0, 3 * f), 9 * H * W,
                                                      3 * H * W, B)))
            for k in range(1, n_ds):
                ops.append((1, lib.b2f_avgpool2x2_forward, (P(ds[k - 1]), P(ds[k]), 2 * B, 3, H >> (k - 1), W >> (k - 1))))
        plan.ds = ds

        # -- training with the tensor-core forward: the (hi, lo) weights follow the flat parameters ---------
        if self.train_planar:
            jobs = [(cv.w.data_ptr(), cv.tc_h.data_ptr(), cv.tc_l.data_ptr(), cv.cout, cv.cin, cv.tc_cin, 0)
                    for cv in self._convs.values() if cv.tc_h is not None]
            table = torch.frombuffer(bytearray(_lib.pack_jobs(jobs)), dtype=torch.uint8).to(dev)
            plan.keep.append(table)
            ops.append((0, lib.b2f_conv3x3_tc_pack_from_packed_batch, (P(table), len(jobs))))       # one launch for all

        # -- siamese feature pyramid on the 3B batch (slots: past, future, reference) -----------------------------
        feats = {}
        slot_of_frame = (0, 2, 1)     # frame index 0, 1, 2 (past, ref, future) -> slot
        prev = None
        for l in range(2, levels + 1):
            h, w = hw(l)
            c_in, c_out = FEAT[l - 2], FEAT[l - 1]
            tmp = E(3 * B, c_out, h, w)
            feats[l] = E(3 * B, c_out, h, w)
            if l == 2:
                for f in range(3):
                    conv("feat.l2.0", sl(x, 0, 3 * f), 9 * H * W, B, 3, H, W, sl(tmp, slot_of_frame[f] * B), 0)
            else:
                conv("feat.l%d.0" % l, P(prev), 0, 3 * B, c_in, 2 * h, 2 * w, P(tmp), 0)
            name = "feat.l%d.1" % l
            if self.tensor_cores and c_out >= TC_FEAT_MIN:
                # the stride-1 half of the convUnit on tcgen05 (levels 4-7; at 16 / 32 channels the M = 128 tensor-core
                # tiles are mostly padding and the FFMA2 kernel is as fast): one split of the stride-2 layer's output, one call for the
                # three frames (192 channels: two output-channel slices; 16: one 32-wide slice), planar output
                cv = self._convs[name]
                th, tl = E(3 * B, h, w, _round32(c_out)), E(3 * B, h, w, _round32(c_out))
                plan.keep += [th, tl]
                plan.tmp_hl[l] = (th, tl)
                ops.append((0, lib.b2f_nhwc_split_from_bdhw, (P(tmp), 0, P(th), P(tl), 3 * B, c_out, h, w)))
                ops.append((0, lib.b2f_conv3x3_tc_forward, (P(th), P(tl), P(cv.tc_h), P(cv.tc_l), P(cv.b), None, None,
                                                            P(feats[l]), 0, 3 * B, c_out, h, w, c_out, C.c_float(0.2))))
                if l >= l_st:     # the reference frame's features also go into the joined decoder input
                    ops.append((0, lib.b2f_copy2d_async, (sl(J[l], 0, 2 * nd), J[l].stride(0), sl(feats[l], 2 * B),
                                                          c_out * h * w, c_out * h * w, B)))
            else:
                conv(name, P(tmp), 0, 2 * B, c_out, h, w, P(feats[l]), 0)
                if l >= l_st:     # the reference frame's features also go into the joined decoder input
                    conv(name, sl(tmp, 2 * B), 0, B, c_out, h, w, sl(feats[l], 2 * B), 0, sl(J[l], 0, 2 * nd), J[l].stride(0))
                else:
                    conv(name, sl(tmp, 2 * B), 0, B, c_out, h, w, sl(feats[l], 2 * B), 0)
            plan.keep.append(tmp)
            plan.tmp[l] = tmp
            prev = feats[l]
        plan.feats = feats

        # -- levels, coarse to fine ---------------------------------------------------------------------------
        outs = {}
        warped = {}
        ufs = {}
        for l in range(levels, l_st - 1, -1):
            h, w = hw(l)
            Cl = FEAT[l - 1]
            ref = sl(feats[l], 2 * B)
            past = sl(feats[l], 0) if l == levels else sl(warped[l], 0)
            fut = sl(feats[l], B) if l == levels else sl(warped[l], B)
            Jl = J[l]
            jbs = Jl.stride(0)
            for fwd, frame, c0 in ((1, fut, 0), (0, past, nd)):
                ops.append((0, lib.b2f_costvol_forward, (_lib.ptr_array([ref.value, frame.value]), 2, B, Cl, h, w, win, fwd,
                                                         sl(Jl, 0, c0), jbs)))
            ops.append(("fork", 2))      # lane 2 (occlusion decoder) may start: J[l] is complete
            cj = Jl.shape[1]

            def decoder(kind, lane, x0, cin0):
                if self.tensor_cores:
                    return decoder_tc(kind, lane, cin0)
                t, tb, cin = x0, jbs, cin0
                chain = []
                for i, cout in enumerate(DEC):
                    out = E(B, cout, h, w)
                    chain.append(out)
                    conv("%s.l%d.%d" % (kind, l, i), t, tb, B, cin, h, w, P(out), 0, slope=0.2 if i < 5 else 1.0, lane=lane)
                    t, tb, cin = P(out), 0, cout
                plan.dec[(kind, l)] = (chain, cin0)
                return out

            def decoder_tc(kind, lane, cin0):
                """Six tensor-core layers over channel-minor (hi, lo) pairs: five wide ones that hand (hi, lo) to the
                next, and the 2-channel head as one 32-wide output slice with two valid channels (the TMA zero-fills
                the weight rows beyond Cout) that writes planar fp32.  No planar copy of the hidden activations: the
                backward plan reads the (hi, lo) inputs of every layer (plan.dec_hl)."""
                xh, xl, cin = Jsplit[0], Jsplit[1], cj
                hl_in = []                         # (hi, lo, channels) of every layer's INPUT
                for i, cout in enumerate(DEC[:5]):
                    hl_in.append((xh, xl, cin))
                    cv = self._convs["%s.l%d.%d" % (kind, l, i)]
                    assert cv.tc_cin == cin, (kind, l, i, cv.tc_cin, cin)
                    oh, ol = E(B, h, w, _round32(cout)), E(B, h, w, _round32(cout))
                    plan.keep += [oh, ol]
                    ops.append((lane, lib.b2f_conv3x3_tc_forward,
                                (P(xh), P(xl), P(cv.tc_h), P(cv.tc_l), P(cv.b), P(oh), P(ol), None, 0, B, cin, h, w, cout,
                                 C.c_float(0.2))))
                    xh, xl, cin = oh, ol, cout
                hl_in.append((xh, xl, cin))
                out = E(B, 2, h, w)
                plan.keep.append(out)
                cv = self._convs["%s.l%d.5" % (kind, l)]
                ops.append((lane, lib.b2f_conv3x3_tc_forward,
                            (P(xh), P(xl), P(cv.tc_h), P(cv.tc_l), P(cv.b), None, None, P(out), 0, B, cin, h, w, 2,
                             C.c_float(1.0))))
                if self.train_planar:
                    plan.dec[(kind, l)] = ([None] * 5 + [out], cin0)
                    plan.dec_hl[(kind, l)] = hl_in
                return out

            Jsplit = None
            if self.tensor_cores:
                cjp = _round32(cj)
                Jsplit = (E(B, h, w, cjp), E(B, h, w, cjp))
                plan.keep += list(Jsplit)
                ops.append((0, lib.b2f_nhwc_split_from_bdhw, (P(Jl), jbs, P(Jsplit[0]), P(Jsplit[1]), B, cj, h, w)))
                ops.append(("fork", 2))

            # occlusion decoder -> softmax -> nearest x 2^(l_st-1) (pwc.lua:288-317)
            occ_logit = decoder("occ", 2, P(Jl), cj)
            occ = E(B, 2, h, w)
            ops.append((2, lib.b2f_softmax_channels_forward, (P(occ_logit), P(occ), B, 2, h, w)))
            up = 1 << (l_st - 1)
            skip_occ = E(B, 2, h * up, w * up)
            ops.append((2, lib.b2f_upsample_nearest_forward, (P(occ), P(skip_occ), B, 2, h, w, up)))
            plan.occ[l] = occ
            plan.skip_occ[l] = skip_occ

            # flow decoder(s) (pwc.lua:322-349)
            flows = []
            for kind in ("flow",) + (("bflow",) if self.past_flow else ()):
                f = decoder(kind, 0, P(Jl), 2 * nd if l == levels else cj)
                flows.append(f)
            plan.fs[l] = flows

            # up-sampling: ufs[l] (next level's decoder input + feature warps), skip_ufs[l] (output + image warps)
            ufs[l] = []
            skips = []
            for i, f in enumerate(flows):
                u = E(B, 2, 2 * h, 2 * w)
                dst = [u.data_ptr()]
                dbs = [0]
                if l > l_st:
                    Jn = J[l - 1]
                    dst.append(Jn.data_ptr() + 4 * (2 * nd + FEAT[l - 2] + 2 * i) * Jn.stride(1))
                    dbs.append(Jn.stride(0))
                ops.append((0, lib.b2f_upsample_bilinear2x_forward,
                            (P(f), 0, B, 2, h, w, _lib.ptr_array(dst), (C.c_int64 * len(dbs))(*dbs), len(dst),
                             C.c_float(1.0))))
                ufs[l].append(u)
            if l > l_st:      # feature warps for the next level (pwc.lua:394-409); always with the FUTURE flow
                hn, wn = hw(l - 1)
                Cn = FEAT[l - 2]
                warped[l - 1] = E(2 * B, Cn, hn, wn)
                for slot, sgn in ((0, -1.0), (1, 1.0)):
                    sc = o.flownet_factor * sgn / 2.0 ** (l - 2)
                    ops.append((0, lib.b2f_warp_bdhw_forward, (sl(feats[l - 1], slot * B), P(ufs[l][0]), C.c_float(sc),
                                                               sl(warped[l - 1], slot * B), B, Cn, hn, wn)))
            ops.append(("fork", 1))      # lane 1: output-resolution up-sampling + image warps, off the critical path
            chains = []
            for u in ufs[l]:
                t = u
                chain = [u]
                for i in range(2, l_st):
                    hh, ww = t.shape[2], t.shape[3]
                    t2 = E(B, 2, 2 * hh, 2 * ww)
                    ops.append((1, lib.b2f_upsample_bilinear2x_forward,
                                (P(t), 0, B, 2, hh, ww, _lib.ptr_array([t2.data_ptr()]), (C.c_int64 * 1)(0), 1,
                                 C.c_float(1.0))))
                    t = t2
                    chain.append(t)
                skips.append(t)
                chains.append(chain)
            plan.skip_chain[l] = chains
            unit = list(skips) + [skip_occ]
            if self.image_warps:
                k = l - l_st
                hh, ww = H >> k, W >> k
                iw = E(2 * B, 3, hh, ww)
                for slot, sgn in ((0, -1.0), (1, 1.0)):
                    fl = skips[1] if (self.past_flow and slot == 0) else skips[0]      # pwc.lua:426-437
                    sc = o.flownet_factor * sgn / 2.0 ** (l - l_st)                    # :443
                    ops.append((1, lib.b2f_warp_bdhw_forward, (sl(ds[k], slot * B), P(fl), C.c_float(sc), sl(iw, slot * B),
                                                               B, 3, hh, ww)))
                unit += [iw[:B], iw[B:]]
                plan.iw[l] = iw
            outs[l] = unit
        plan.ufs = ufs
        plan.warped = warped
        plan.output = [t for l in range(l_st, levels + 1) for t in outs[l]]
        return plan

    # ---- backward plan ------------------------------------------------------------------------------------
    def _ensure_training_state(self):
        """Gradient buffer (same flat layout as the parameters), transposed weights for backward-data, Adam moments."""
        if getattr(self, "flat_grads", None) is not None:
            return
        lib, dev = self.lib, self.device
        n = self.flat_params.numel()
        self.flat_grads = torch.empty(n, device=dev, dtype=torch.float32)
        self.adam_m = torch.empty(n, device=dev, dtype=torch.float32)
        self.adam_v = torch.empty(n, device=dev, dtype=torch.float32)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for t in (self.flat_grads, self.adam_m, self.adam_v):
            _lib.check(lib.b2f_zero_async(_vp(t), n * 4, st))
        self.adam_t = 0
        for name, cv in self._convs.items():
            nw = cv.b_off - cv.w_off
            cv.gw = self.flat_grads[cv.w_off:cv.w_off + nw]
            cv.gb = self.flat_grads[cv.b_off:cv.b_off + cv.cout]
            cv.wt = torch.empty(int(lib.b2f_conv3x3_packed_floats(cv.cout, cv.cin)), device=dev, dtype=torch.float32)

    def _build_backward(self, plan):
        """`model:backward(input, gradOutputs)` (train.lua:480) as a list of C-ABI calls: the forward plan walked in
        reverse.  Gradients with several consumers (nngraph's fan-out nodes) are accumulated with b2f_axpy2d or with
        the accumulate flag of backward-data; the LeakyReLU derivative of a layer is applied by the backward-data
        call that produces its output gradient (decoders) or once on the summed gradient (feature pyramid)."""
        self._ensure_training_state()
        o, lib, dev = self.opt, self.lib, self.device
        B, H, W = plan.shape
        levels, l_st, win = o.levels, o.l_st, o.pwc_ws
        nd = win * win
        E = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        hw = lambda l: (H >> (l - 1), W >> (l - 1))
        P = _vp
        ops = plan.bops = []
        nflow = 2 if self.past_flow else 1

        def sl(t, b0=0, c0=0):
            return C.c_void_p(t.data_ptr() + 4 * (b0 * t.stride(0) + c0 * t.stride(1)))

        def zero(t):
            ops.append((lib.b2f_zero_async, (P(t), t.numel() * 4)))

        def axpy(dst, dstride, src, sstride, row, rows, alpha=1.0):
            ops.append((lib.b2f_axpy2d, (dst, dstride or row, src, sstride or row, row, rows, C.c_float(alpha))))

        def wgrad(name, xin, xbs, gout, gbs, nb, cin, h, w):
            cv = self._convs[name]
            assert cv.cin == cin
            ops.append((lib.b2f_conv3x3_backward_weights, (xin, xbs, gout, gbs, P(cv.gw), P(cv.gb), nb, cin, h, w, cv.cout,
                                                           cv.stride), 1))

        def dgrad(name, gout, gbs, act, abs_, gin, ibs, acc, nb, cin, h, w, slope=0.2, stride=None):
            cv = self._convs[name]
            ops.append((lib.b2f_conv3x3_backward_data, (gout, gbs, P(cv.wt), act, abs_, gin, ibs, int(acc), nb, cin, h, w,
                                                        cv.cout, stride or cv.stride, C.c_float(slope))))

        def dec_range(l):
            """[lo, hi) of level l's decoders in the flat parameter / gradient buffer (they are consecutive there)."""
            kinds = ("occ", "flow") + (("bflow",) if self.past_flow else ())
            first, last = self._convs["%s.l%d.0" % (kinds[0], l)], self._convs["%s.l%d.5" % (kinds[-1], l)]
            return first.w_off, last.b_off + _round64(last.cout)

        # Gradient buckets for the data-parallel all-reduce: (index into bops AFTER which the flat range
        # [lo, hi) is final).  Backward finishes the finest level's decoders first and the first convUnit last.
        plan.bucket_marks = []
        # gradOutputs, in the order of the output table
        plan.gout = [E(*t.shape) for t in plan.output]
        per = len(plan.output) // (levels - l_st + 1)
        zero(self.flat_grads)                                      # model:zeroGradParameters(), train.lua:251
        jobs = []
        for name, cv in self._convs.items():
            kind, _lvl, idx = name.split(".")
            if self.train_planar and (kind in ("occ", "flow", "bflow") or
                                      (kind == "feat" and idx == "1" and cv.cout >= TC_FEAT_MIN) or
                                      (kind == "feat" and idx == "0" and cv.cin > 3)):
                # input gradient on tcgen05: transposed, mirrored (hi, lo) operand; all of them in ONE launch below
                if cv.tct_h is None:
                    n_t = 9 * cv.cin * _round32(cv.cout)
                    cv.tct_h = torch.empty(n_t, device=dev, dtype=torch.float32)
                    cv.tct_l = torch.empty(n_t, device=dev, dtype=torch.float32)
                jobs.append((cv.w.data_ptr(), cv.tct_h.data_ptr(), cv.tct_l.data_ptr(), cv.cout, cv.cin, cv.cout, 1))
            elif not (kind == "feat" and idx == "0" and cv.cin == 3):
                # input gradient on the FFMA kernel (the first convolution's input is the image: no input gradient)
                ops.append((lib.b2f_conv3x3_transpose_packed, (P(cv.w), P(cv.wt), cv.cout, cv.cin)))
        if jobs:
            table = torch.frombuffer(bytearray(_lib.pack_jobs(jobs)), dtype=torch.uint8).to(dev)
            plan.keep.append(table)
            ops.append((lib.b2f_conv3x3_tc_pack_from_packed_batch, (P(table), len(jobs))))
        g_feats = {l: E(*plan.feats[l].shape) for l in plan.feats}
        for l in g_feats:
            zero(g_feats[l])
        gJ = {l: E(*plan.J[l].shape) for l in plan.J}
        g_warped = {l: E(*plan.warped[l].shape) for l in plan.warped}
        plan.g_feats, plan.gJ = g_feats, gJ

        for l in range(l_st, levels + 1):
            h, w = hw(l)
            k = l - l_st
            Cl = FEAT[l - 1]
            unit = plan.gout[k * per:(k + 1) * per]
            g_skip = unit[:nflow]
            g_skip_occ = unit[nflow]
            hh, ww = g_skip[0].shape[2], g_skip[0].shape[3]
            # image warps: flow gradient only (the frames are inputs), pwc.lua:441-446
            if self.image_warps:
                tmpf = E(B, 2, hh, ww)
                for slot, sgn in ((0, -1.0), (1, 1.0)):
                    fi = 1 if (self.past_flow and slot == 0) else 0
                    sc = o.flownet_factor * sgn / 2.0 ** (l - l_st)
                    ops.append((lib.b2f_warp_bdhw_backward, (sl(plan.ds[k], slot * B), P(plan.skip_chain[l][fi][-1]),
                                                             C.c_float(sc), P(unit[nflow + 1 + slot]), None, P(tmpf),
                                                             B, 3, hh, ww)))
                    axpy(P(g_skip[fi]), 0, P(tmpf), 0, tmpf.numel(), 1)
                plan.keep.append(tmpf)
            g_u = []
            for fi in range(nflow):
                g = g_skip[fi]
                chain = plan.skip_chain[l][fi]
                for t in reversed(chain[:-1]):                   # output-resolution up-samplings, :374-389
                    g2 = E(*t.shape)
                    ops.append((lib.b2f_upsample_bilinear2x_backward, (P(g), P(g2), B, 2, t.shape[2], t.shape[3],
                                                                       C.c_float(1.0), 0)))
                    g = g2
                if g is g_skip[fi]:                               # l_st == 2: the up-sampled flow IS the output
                    g2 = E(*g.shape)
                    ops.append((lib.b2f_copy2d_async, (P(g2), g.numel(), P(g), g.numel(), g.numel(), 1)))
                    g = g2
                g_u.append(g)
                plan.keep.append(g)         # every buffer an op points at must outlive this function
                if l > l_st:                                      # the next level's decoders read ufs[l] from J[l-1]
                    Jn = gJ[l - 1]
                    axpy(P(g), 0, sl(Jn, 0, 2 * nd + FEAT[l - 2] + 2 * fi), Jn.stride(0), g.numel() // B, B)
            if l > l_st:                                          # feature warps, :402-408 (future flow only)
                hn, wn = hw(l - 1)
                Cn = FEAT[l - 2]
                tmpf = E(B, 2, hn, wn)
                for slot, sgn in ((0, -1.0), (1, 1.0)):
                    sc = o.flownet_factor * sgn / 2.0 ** (l - 2)
                    ops.append((lib.b2f_warp_bdhw_backward, (sl(plan.feats[l - 1], slot * B), P(plan.ufs[l][0]), C.c_float(sc),
                                                             sl(g_warped[l - 1], slot * B), sl(g_feats[l - 1], slot * B),
                                                             P(tmpf), B, Cn, hn, wn)))
                    axpy(P(g_u[0]), 0, P(tmpf), 0, tmpf.numel(), 1)
                plan.keep.append(tmpf)
            Jl, gJl = plan.J[l], gJ[l]
            jbs = Jl.stride(0)

            def decoder_backward(kind, G, first, lane=0):
                chain, cin0 = plan.dec[(kind, l)]
                if self.train_planar:
                    return decoder_backward_tc(kind, G, first, chain, cin0, lane)
                for i in range(5, -1, -1):
                    name = "%s.l%d.%d" % (kind, l, i)
                    if i > 0:
                        xin, xbs, cin = P(chain[i - 1]), 0, DEC[i - 1]
                    else:
                        xin, xbs, cin = P(Jl), jbs, cin0
                    wgrad(name, xin, xbs, P(G), 0, B, cin, h, w)
                    if i > 0:
                        gin = E(B, cin, h, w)
                        plan.keep.append(gin)
                        dgrad(name, P(G), 0, P(chain[i - 1]), 0, P(gin), 0, False, B, cin, h, w)
                        G = gin
                    else:
                        dgrad(name, P(G), 0, None, 0, P(gJl), jbs, not first, B, cin, h, w)

            def decoder_backward_tc(kind, G, first, chain, cin0, lane):
                """The same walk on tcgen05: the 2-channel output gradient is split once into channel-minor (hi, lo);
                the input gradients of layers 5..1 (b2f_conv3x3_tc_backward_data: the forward tensor-core kernel on
                transposed, mirrored (hi, lo) weights, LeakyReLU derivative from the hi half of the layer's input) hand
                (hi, lo) to the next layer; the WEIGHT gradients of layers 5..0 (b2f_conv3x3_tc_backward_weights:
                MN-major operands, contraction over pixels) read the (hi, lo) input activations the forward left
                behind and the (hi, lo) output gradients, the bias gradients are summed from (hi, lo) too.  Layer 0's
                input gradient (162 .. 356 channels) runs as slices of <= 128 channels into the joined gradient."""
                hl_in = plan.dec_hl[(kind, l)]
                gh, gl = E(B, h, w, 32), E(B, h, w, 32)
                # `lane` != 0: this chain runs beside the main stream's; it starts behind what produced G there
                ops.append((lib.b2f_nhwc_split_from_bdhw, (P(G), 0, P(gh), P(gl), B, 2, h, w), lane, (0,)))
                plan.keep += [gh, gl]
                Gp = G
                for i in range(5, -1, -1):
                    name = "%s.l%d.%d" % (kind, l, i)
                    cv = self._convs[name]
                    xh, xl, cx = hl_in[i]
                    ops.append((lib.b2f_conv3x3_tc_backward_weights,
                                (P(xh), P(xl), cx, P(gh), P(gl), P(Gp) if Gp is not None else None, 0, P(cv.gw), P(cv.gb), B,
                                 cv.cin, h, w, cv.cout), 1, (lane,)))
                    if i == 0:
                        break
                    cin, cout = DEC[i - 1], cv.cout
                    nh, nl = E(B, h, w, _round32(cin)), E(B, h, w, _round32(cin))
                    plan.keep += [nh, nl]
                    # the LeakyReLU derivative from the HI half of layer i's channel-minor input (= layer i - 1's output);
                    # no planar copy of the gradient: the next weight-gradient call sums the bias gradient from (hi, lo)
                    ops.append((lib.b2f_conv3x3_tc_backward_data,
                                (P(gh), P(gl), P(cv.tct_h), P(cv.tct_l), None, 0, P(xh), P(nh), P(nl), None, 0, B, cout,
                                 h, w, cin, C.c_float(0.2), 0), lane))
                    gh, gl, Gp = nh, nl, None
                # layer 0: 162 .. 356 input channels as slices of <= 128, straight into (or added to) the joined gradient
                cv0 = self._convs["%s.l%d.0" % (kind, l)]
                assert cv0.cin == cin0
                last = (lib.b2f_conv3x3_tc_backward_data,
                        (P(gh), P(gl), P(cv0.tct_h), P(cv0.tct_l), None, 0, None, None, None, P(gJl), jbs, B, cv0.cout, h,
                         w, cin0, C.c_float(1.0), 0 if first else 1))
                if lane == 0:
                    ops.append(last)
                    return None
                return last         # the caller places it behind the occlusion decoder's write of the joined gradient

            # occlusion path: nearest^T, softmax^T, decoder (first: it reads every channel of J[l])
            g_occ = E(B, 2, h, w)
            ops.append((lib.b2f_upsample_nearest_backward, (P(g_skip_occ), P(g_occ), B, 2, h, w, 1 << (l_st - 1))))
            g_logit = E(B, 2, h, w)
            ops.append((lib.b2f_softmax_channels_backward, (P(plan.occ[l]), P(g_occ), P(g_logit), B, 2, h, w)))
            plan.keep += [g_occ, g_logit]
            g_fs = [E(B, 2, h, w) for _ in range(nflow)]
            plan.keep += g_fs
            for fi in range(nflow):
                ops.append((lib.b2f_upsample_bilinear2x_backward, (P(g_u[fi]), P(g_fs[fi]), B, 2, h, w, C.c_float(1.0), 0)))
            if self.train_planar:
                # flow decoders' chains on lanes 2 / 3 beside the occlusion decoder's on the main stream; their last
                # calls ADD into the joined gradient: behind the occlusion decoder's write, and one after the other
                # (two concurrent reductions into the same words would make the sum's rounding order a race)
                lasts = [decoder_backward(kind, g_fs[fi], False, lane=2 + fi) for fi, kind in enumerate(("flow", "bflow")[:nflow])]
                decoder_backward("occ", g_logit, True)
                for fi, last in enumerate(lasts):
                    ops.append((last[0], last[1], 2 + fi, (0,) + tuple(range(2, 2 + fi))))
            else:
                decoder_backward("occ", g_logit, True)
                for fi, kind in enumerate(("flow", "bflow")[:nflow]):
                    decoder_backward(kind, g_fs[fi], False)
            # cost volumes: gradRef of both directions + the joined input's feature slice -> reference features
            ref = sl(plan.feats[l], 2 * B)
            tmp_ref, tmp_frm = E(B, Cl, h, w), E(B, Cl, h, w)
            plan.keep += [tmp_ref, tmp_frm]
            n_item = Cl * h * w
            for fwd, slot, c0 in ((1, 1, 0), (0, 0, nd)):
                frame = sl(plan.feats[l], slot * B) if l == levels else sl(plan.warped[l], slot * B)
                gfrm = P(tmp_frm) if l == levels else sl(g_warped[l], slot * B)
                ops.append((lib.b2f_costvol_backward, (_lib.ptr_array([ref.value, frame.value]), 2, B, Cl, h, w, win, fwd,
                                                       sl(gJl, 0, c0), jbs, _lib.ptr_array([tmp_ref.data_ptr(), gfrm.value])),
                            0, (2, 3)))
                axpy(sl(g_feats[l], 2 * B), 0, P(tmp_ref), 0, B * n_item, 1)
                if l == levels:
                    axpy(sl(g_feats[l], slot * B), 0, P(tmp_frm), 0, B * n_item, 1)
            axpy(sl(g_feats[l], 2 * B), n_item, sl(gJl, 0, 2 * nd), jbs, n_item, B)
            plan.bucket_marks.append((len(ops), dec_range(l)))

        # feature pyramid, coarse to fine (pwc.lua:176-211); the three frames share the weights: one batch of 3B
        for l in range(levels, 1, -1):
            h, w = hw(l)
            c_in, c_out = FEAT[l - 2], FEAT[l - 1]
            gf, f, tmp = g_feats[l], plan.feats[l], plan.tmp[l]
            g_tmp_hl = None
            ops.append((lib.b2f_leaky_relu_backward, (P(gf), gf.numel(), P(f), f.numel(), gf.numel(), 1, C.c_float(0.2))))
            if self.train_planar and c_out >= TC_FEAT_MIN:
                cv1 = self._convs["feat.l%d.1" % l]
                th, tl = plan.tmp_hl[l]
                gfh, gfl = E(3 * B, h, w, _round32(c_out)), E(3 * B, h, w, _round32(c_out))
                g_tmp = E(*tmp.shape)
                plan.keep += [gfh, gfl, g_tmp]
                ops.append((lib.b2f_nhwc_split_from_bdhw, (P(gf), 0, P(gfh), P(gfl), 3 * B, c_out, h, w)))
                ops.append((lib.b2f_conv3x3_tc_backward_weights,
                            (P(th), P(tl), c_out, P(gfh), P(gfl), P(gf), 0, P(cv1.gw), P(cv1.gb), 3 * B, c_out, h, w, c_out), 1))
                if c_out in (64, 96, 128):      # one slice: the (hi, lo) form for the stride-2 input gradient below comes for free
                    gth, gtl = E(3 * B, h, w, c_out), E(3 * B, h, w, c_out)
                    plan.keep += [gth, gtl]
                    g_tmp_hl = (gth, gtl)
                ops.append((lib.b2f_conv3x3_tc_backward_data,
                            (P(gfh), P(gfl), P(cv1.tct_h), P(cv1.tct_l), None, 0, P(th), P(g_tmp_hl[0]) if g_tmp_hl else None,
                             P(g_tmp_hl[1]) if g_tmp_hl else None, P(g_tmp), 0, 3 * B, c_out, h, w, c_out, C.c_float(0.2), 0)))
            else:
                wgrad("feat.l%d.1" % l, P(tmp), 0, P(gf), 0, 3 * B, c_out, h, w)
                g_tmp = E(*tmp.shape)
                plan.keep.append(g_tmp)
                dgrad("feat.l%d.1" % l, P(gf), 0, P(tmp), 0, P(g_tmp), 0, False, 3 * B, c_out, h, w)
            name = "feat.l%d.0" % l
            if l == 2:
                for fr, slot in enumerate((0, 2, 1)):
                    wgrad(name, sl(plan.x, 0, 3 * fr), 9 * H * W, sl(g_tmp, slot * B), 0, B, 3, H, W)
            else:
                wgrad(name, P(plan.feats[l - 1]), 0, P(g_tmp), 0, 3 * B, c_in, 2 * h, 2 * w)
                if self.train_planar:
                    # stride-2 input gradient on tcgen05: four parity-class accumulators over the low-resolution gradient
                    cv0 = self._convs[name]
                    if g_tmp_hl is None:
                        g_tmp_hl = (E(3 * B, h, w, _round32(c_out)), E(3 * B, h, w, _round32(c_out)))
                        plan.keep += list(g_tmp_hl)
                        ops.append((lib.b2f_nhwc_split_from_bdhw, (P(g_tmp), 0, P(g_tmp_hl[0]), P(g_tmp_hl[1]), 3 * B, c_out, h, w)))
                    ops.append((lib.b2f_conv3x3_tc_backward_data_s2,
                                (P(g_tmp_hl[0]), P(g_tmp_hl[1]), P(cv0.tct_h), P(cv0.tct_l), P(g_feats[l - 1]), 0, 3 * B, c_out, h,
                                 w, c_in, 2 * h, 2 * w, 1)))
                elif w % 2 == 0:
                    # stride 2: dilate the output gradient and run the stride-1 (TMA / FFMA2) input-gradient kernel
                    z = E(3 * B, c_out, 2 * h, 2 * w)
                    plan.keep.append(z)
                    ops.append((lib.b2f_zero_insert2x, (P(g_tmp), P(z), 3 * B * c_out, h, w)))
                    dgrad(name, P(z), 0, None, 0, P(g_feats[l - 1]), 0, True, 3 * B, c_in, 2 * h, 2 * w, slope=1.0, stride=1)
                else:
                    dgrad(name, P(g_tmp), 0, None, 0, P(g_feats[l - 1]), 0, True, 3 * B, c_in, 2 * h, 2 * w, slope=1.0)
            c0_, c1_ = self._convs["feat.l%d.0" % l], self._convs["feat.l%d.1" % l]
            plan.bucket_marks.append((len(ops), (c0_.w_off, c1_.b_off + _round64(c1_.cout))))
        plan.g_keep = (g_warped,)

    def backward(self, x, gradOutputs, graph=False):
        """model:backward(x, gradOutputs) after a forward of the same x: accumulates d loss / d parameters into
        `flat_grads` (zeroed first, like train.lua:251's zeroGradParameters).  gradOutputs: tensors shaped like the
        output table (device)."""
        p = self.plan(x.size(0), x.size(2), x.size(3))
        if not self.image_warps or (self.tensor_cores and not self.train_planar):
            raise RuntimeError("backward needs the full output table and the planar activations: build the model with "
                               "image_warps=True and tensor_cores=False (or tensor_cores=True, train_planar=True)")
        with torch.cuda.device(self.device):
            if p.bops is None:
                self._build_backward(p)
            if len(gradOutputs) != len(p.gout):
                raise ValueError("backward: %d gradOutputs for %d outputs" % (len(gradOutputs), len(p.gout)))
            for dst, src in zip(p.gout, gradOutputs):
                if tuple(src.shape) != tuple(dst.shape):
                    raise ValueError("backward: gradOutput of shape %r for an output of shape %r" %
                                     (tuple(src.shape), tuple(dst.shape)))
                dst.copy_(src, non_blocking=True)
            self.run_backward(p, graph=graph)
        return self.flat_grads

    def run_backward(self, p, graph=False):
        if p.bops is None:
            self._build_backward(p)
        if not graph:
            p.launch_backward()
            return
        if p.bgraph is None:
            saved = [g.clone() for g in p.gout]
            p.launch_backward()
            torch.cuda.current_stream().synchronize()
            for g, sv in zip(p.gout, saved):
                g.copy_(sv)
            g = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream(self.device)
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.graph(g, stream=cap):
                p.launch_backward()
            p.bgraph = g
        p.bgraph.replay()

    def grad_params(self):
        """flat_grads back in Torch's layout (dict of numpy arrays, like state_params)."""
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        out = {}
        with torch.cuda.device(self.device):
            for name, cout, cin0, _s in conv_shapes(self.opt):
                cv = self._convs[name]
                w = torch.empty((cout, cv.cin, 3, 3), device=self.device)
                _lib.check(self.lib.b2f_conv3x3_pack_weights(_vp(w), _vp(cv.gw), cout, cv.cin, 1, st))
                w = w.cpu().numpy()
                if cv.cin != cin0:
                    kind = name.split(".")[0]
                    w = np.concatenate([w[:, :cin0 - 2], w[:, cin0:cin0 + 2] if kind == "bflow" else w[:, cin0 - 2:cin0]], 1)
                out[name + ".weight"] = np.ascontiguousarray(w)
                out[name + ".bias"] = cv.gb.cpu().numpy().copy()
        return out

    def adam_step(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        """optim.adam(feval, parameters, optimState) (train.lua:485-486) on the flat buffers."""
        self._ensure_training_state()
        self.adam_t += 1
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.b2f_adam_step(_vp(self.flat_params), _vp(self.flat_grads), _vp(self.adam_m), _vp(self.adam_v),
                                          self.flat_params.numel(), lr, beta1, beta2, eps, weight_decay, self.adam_t, st))

    # ---- execution --------------------------------------------------------------------------------------
    def plan(self, B, H, W, slot=0):
        """The plan (buffers + call list) of one input shape.  `slot` > 0: further, independent sets of buffers for the
        same shape -- two plans replayed on two streams overlap one batch's coarse levels (launches that cannot fill
        the machine) with the other's fine ones (bench.py `whole_network.device_two_streams`)."""
        key = (B, H, W) if slot == 0 else (B, H, W, slot)
        if key not in self._plans:
            with torch.cuda.device(self.device):
                self._plans[key] = self._build(B, H, W)
        return self._plans[key]

    def forward(self, x, graph=True):
        """model:forward(x): x (B, 9, H, W) float32 on this device (or a pinned / pageable HOST tensor: it is copied
        into the plan's input buffer on the current stream).  Returns the output table (views of the plan's buffers,
        valid until the next forward of the same shape)."""
        if x.dim() != 4 or x.size(1) != 9:
            raise ValueError("PWCNet: expected a (B, 9, H, W) input, got %r" % (tuple(x.shape),))
        if x.dtype != torch.float32:
            raise TypeError("PWCNet: expected float32")
        p = self.plan(x.size(0), x.size(2), x.size(3))
        with torch.cuda.device(self.device):
            p.x.copy_(x, non_blocking=True)      # a cudaMemcpyAsync (H2D or D2D), no kernel
            self.run(p, graph=graph)
        self.output = p.output
        return p.output

    def run(self, p, graph=True):
        """Execute the plan on the current stream (inputs already in p.x)."""
        if not graph:
            p.launch()
            return
        if p.graph is None:
            p.launch()                            # warm-up outside capture: TMA descriptors, func attributes
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream(self.device)
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.graph(g, stream=cap):
                p.launch()
            p.graph = g
        p.graph.replay()

    def evaluate(self):
        return self

    def cuda(self):
        return self


class _Plan:
    def __init__(self, B, H, W):
        self.shape = (B, H, W)
        self.ops = []
        self.keep = []
        self.occ = {}
        self.fs = {}
        self.tmp, self.dec, self.skip_occ, self.skip_chain, self.iw = {}, {}, {}, {}, {}
        self.dec_hl = {}
        self.tmp_hl = {}
        self.bops = None
        self.bgraph = None
        self.graph = None
        self._lanes = None
        self._blanes = None
        self._bused = set()
        self.n_launches = 0

    def launch(self):
        """Issue every call.  Lane 0 is the current stream (the critical path: cost volumes -> flow decoder -> up-sampling
        -> feature warps -> next level); lanes 1 (output up-sampling, image warps) and 2 (occlusion decoders) are side
        streams forked from lane 0 by events and joined at the end."""
        cur = torch.cuda.current_stream()
        if self._lanes is None:
            self._lanes = {1: torch.cuda.Stream(), 2: torch.cuda.Stream()}
        lanes = {0: cur, 1: self._lanes[1], 2: self._lanes[2]}
        used = set()
        check = _lib.check
        n = 0
        for op in self.ops:
            if op[0] == "fork":
                ev = torch.cuda.Event()
                ev.record(cur)
                lanes[op[1]].wait_event(ev)
                used.add(op[1])
                continue
            lane, fn, args = op
            if lane and lane not in used:      # work that depends on the input only
                lanes[lane].wait_stream(cur)
                used.add(lane)
            check(fn(*args, C.c_void_p(lanes[lane].cuda_stream)))
            n += 1
        for lane in used:
            cur.wait_stream(lanes[lane])
        self.n_launches = n

    def launch_backward(self, lo=0, hi=None, join=True):
        """Issue bops[lo:hi].  An entry is (fn, args[, lane[, deps]]): lane 0 is the current stream; lane 1 carries the
        weight gradients (leaves of the backward graph that only the all-reduce / Adam read); lanes 2, 3 the input-
        gradient chains of the flow decoders, which are independent of the occlusion decoder's until both add into the
        joined gradient.  `deps` = lanes whose work issued so far the call must wait for (an event recorded at that
        point; default: lane 0 for a weight gradient, nothing otherwise -- a lane is ordered in itself).  All lanes
        are joined into the current stream at the end of the slice unless join=False (the data-parallel step issues
        the plan bucket by bucket and lets only the COMMUNICATION stream wait for the lanes: `backward_streams()`).  Every buffer a side-lane call reads is written
        once per step, so nothing later on another lane can overwrite it; at the coarse levels, where no kernel fills
        the machine, the chains and the weight gradients run under each other."""
        cur = torch.cuda.current_stream()
        if self._blanes is None:
            self._blanes = {k: torch.cuda.Stream() for k in (1, 2, 3)}
        streams = {0: cur}
        streams.update(self._blanes)
        handles = {k: C.c_void_p(v.cuda_stream) for k, v in streams.items()}
        if lo == 0:
            self._bused = set()
        used = self._bused
        check = _lib.check
        for op in self.bops[lo:hi]:
            lane = op[2] if len(op) > 2 else 0
            deps = op[3] if len(op) > 3 else ((0,) if lane == 1 else ())
            if not SIDE_LANES:
                lane, deps = 0, ()
            for d in deps:
                if d != lane and (d == 0 or d in used):
                    ev = torch.cuda.Event()
                    ev.record(streams[d])
                    streams[lane].wait_event(ev)
            check(op[0](*op[1], handles[lane]))
            if lane:
                used.add(lane)
        if join:
            for lane in used:
                cur.wait_stream(streams[lane])
            used.clear()

    def backward_streams(self):
        """The side streams that carry un-joined work of the backward plan (after launch_backward(..., join=False))."""
        return [self._blanes[k] for k in sorted(self._bused)] if self._blanes else []
