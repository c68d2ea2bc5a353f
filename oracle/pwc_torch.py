"""The multi-frame PWC network as a torch CPU float64 autograd graph -- TEST INFRASTRUCTURE, NOT PRODUCT.

Second restatement of models/pwc.lua:139-492 (the first is oracle/pwc_oracle.py, numpy): the convolutional modules
are torch.nn.functional ops (the descendants of the THNN routines Torch7 calls), the two hot-path modules are
autograd Functions around the numpy oracle's forward / backward (oracle/b2f_oracle.py: CostVolMulti.lua:49-181,
BilinearSamplerBHWD.cu:41-115, 161-307).  What it is for: `model:backward(inputs, gradOutputs)` (train.lua:480) of the
reference is, for given gradOutputs, the gradient of sum_k <output_k, gradOutput_k> with respect to the parameters;
autograd of this graph yields exactly that, which pins the hand-written backward plan of back2future_b200/pwc.py.
Only tests/, smoke() and bench.py's checker legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import b2f_oracle as o
from . import pwc_oracle as po


class _CostVol(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, frame, win, fwd):
        ctx.save_for_backward(ref, frame)
        ctx.win, ctx.fwd = win, fwd
        return torch.from_numpy(o.costvol_forward([ref.detach().numpy(), frame.detach().numpy()], win, fwd))

    @staticmethod
    def backward(ctx, go):
        ref, frame = ctx.saved_tensors
        g = o.costvol_backward([ref.detach().numpy(), frame.detach().numpy()], go.numpy(), ctx.win, ctx.fwd)
        return torch.from_numpy(np.asarray(g[0])), torch.from_numpy(np.asarray(g[1])), None, None


class _WarpUnit(torch.autograd.Function):
    """warpingUnit(I, MulConstant(scale)(F)) of pwc.lua:68-73, 404, 443 (BDHW in / out).  The sampler oracle works on
    fp32 inputs (it restates fp32 address arithmetic); its gradients are returned in float64."""

    @staticmethod
    def forward(ctx, img, flow, scale):
        ctx.save_for_backward(img, flow)
        ctx.scale = scale
        return torch.from_numpy(o.warping_unit_forward(img.detach().numpy().astype(np.float32),
                                                       flow.detach().numpy().astype(np.float32), scale))

    @staticmethod
    def backward(ctx, go):
        img, flow = ctx.saved_tensors
        gi, gf = o.warping_unit_backward(img.detach().numpy().astype(np.float32), flow.detach().numpy().astype(np.float32),
                                         ctx.scale, go.numpy())
        return torch.from_numpy(np.ascontiguousarray(gi, dtype=np.float64)), \
            torch.from_numpy(np.ascontiguousarray(gf, dtype=np.float64)), None


def make_params(params):
    """dict of numpy arrays -> dict of float64 leaf tensors with requires_grad."""
    return {k: torch.from_numpy(np.asarray(v, np.float64)).requires_grad_(True) for k, v in params.items()}


def forward(P, x, opt=None):
    """createModelMulti(opt):forward(x) with torch ops; P from make_params, x (B, 9, H, W) array.  Returns the output
    table (list of tensors, finest level first) like pwc_oracle.pwc_forward."""
    opt = opt or po.Opt()
    xt = torch.from_numpy(np.asarray(x, np.float64))
    levels, l_st, win, ref = opt.levels, opt.l_st, opt.pwc_ws, 2

    def unit(name, t):
        t = F.leaky_relu(F.conv2d(t, P[name + ".0.weight"], P[name + ".0.bias"], stride=2, padding=1), 0.2)
        return F.leaky_relu(F.conv2d(t, P[name + ".1.weight"], P[name + ".1.bias"], padding=1), 0.2)

    def dec(name, t):
        for i in range(6):
            t = F.conv2d(t, P["%s.%d.weight" % (name, i)], P["%s.%d.bias" % (name, i)], padding=1)
            if i < 5:
                t = F.leaky_relu(t, 0.2)
        return t

    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    Is = {f: xt[:, 3 * (f - 1):3 * f] for f in (1, 2, 3)}
    cs = {}
    for f in (1, 2, 3):
        pyr = [Is[f]]
        for l in range(2, levels + 1):
            pyr.append(unit("feat.l%d" % l, pyr[-1]))
        cs[f] = pyr
    ds = {}
    for f in (1, 3):
        ds[f] = [Is[f]]
        for _ in range(2, levels - l_st + 2):
            ds[f].append(F.avg_pool2d(ds[f][-1], 2, 2))
    ws = {1: {}, 3: {}}
    ufs, ubfs, outs = {}, {}, {}
    for l in range(levels, l_st - 1, -1):
        refl = cs[ref][l - 1]
        fut = cs[3][l - 1] if l == levels else ws[3][l]
        past = cs[1][l - 1] if l == levels else ws[1][l]
        cvs = torch.cat([_CostVol.apply(refl, fut, win, True), _CostVol.apply(refl, past, win, False)], 1)
        occ_in = [cvs, refl] + ([ufs[l + 1]] if l != levels else [])
        occ = F.softmax(dec("occ.l%d" % l, torch.cat(occ_in, 1)), dim=1)
        so = occ
        for _ in range(1, l_st):
            so = F.interpolate(so, scale_factor=2, mode="nearest")
        if l == levels:
            fl = dec("flow.l%d" % l, cvs)
            bfl = dec("bflow.l%d" % l, cvs) if opt.past_flow else None
        else:
            fl = dec("flow.l%d" % l, torch.cat([cvs, refl, ufs[l + 1]], 1))
            bfl = dec("bflow.l%d" % l, torch.cat([cvs, refl, ubfs[l + 1]], 1)) if opt.past_flow else None
        ufs[l] = up(fl)
        su = ufs[l]
        for _ in range(2, l_st):
            su = up(su)
        sb = None
        if opt.past_flow:
            ubfs[l] = up(bfl)
            sb = ubfs[l]
            for _ in range(2, l_st):
                sb = up(sb)
        unit_out = [su] + ([sb] if opt.past_flow else []) + [so]
        for f in (1, 3):
            if l > l_st:
                ws[f][l - 1] = _WarpUnit.apply(cs[f][l - 2], ufs[l], opt.flownet_factor * (f - ref) / 2.0 ** (l - 2))
            tmp = sb if (opt.past_flow and f < ref) else su
            unit_out.append(_WarpUnit.apply(ds[f][l - l_st], tmp, opt.flownet_factor * (f - ref) / 2.0 ** (l - l_st)))
        outs[l] = unit_out
    return [t for l in range(l_st, levels + 1) for t in outs[l]]


def backward(params, x, grad_outputs, opt=None):
    """model:backward(x, gradOutputs): (outputs, dict name -> d sum_k <out_k, gradOut_k> / d param), float64 numpy."""
    P = make_params(params)
    outs = forward(P, x, opt)
    assert len(outs) == len(grad_outputs)
    total = sum((t * torch.from_numpy(np.asarray(g, np.float64))).sum() for t, g in zip(outs, grad_outputs))
    total.backward()
    return [t.detach().numpy() for t in outs], {k: v.grad.numpy() for k, v in P.items()}
