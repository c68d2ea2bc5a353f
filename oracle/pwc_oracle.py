"""CPU restatement of the conv trunk and of the whole multi-frame PWC network -- TEST INFRASTRUCTURE, NOT PRODUCT.

Follows models/pwc.lua (reference root): convUnit :58-65, warpingUnit :68-73, decoder :76-85, createModelMulti
:87-508.  The per-module arithmetic is Torch7's nn / THNN (external to the reference tree), restated
from its published definitions:

  nn.SpatialConvolution(nIn, nOut, 3, 3, s, s, 1, 1)  cross-correlation, zero padding 1, bias added
  nn.LeakyReLU(0.2)                                   x > 0 ? x : 0.2 x
  nn.SpatialAveragePooling(2, 2, 2, 2)                mean of 2 x 2 blocks, floor mode
  nn.SpatialUpSamplingBilinear(2)                     align-corners mapping src = dst (in - 1) / (out - 1)
  nn.SpatialUpSamplingNearest(2)                      out[y, x] = in[y // 2, x // 2]
  nn.SpatialSoftMax                                   softmax over the channel dimension
  nn.JoinTable(2), nn.MulConstant, nn.Narrow          layout only

Pinned (tests/test_pwc_oracle.py) against torch's CPU float64 functional ops -- conv2d, avg_pool2d,
interpolate(bilinear, align_corners=True), interpolate(nearest), softmax -- the descendants of the THNN routines
Torch7 calls; the graph wiring itself has no executable reference here (no LuaJIT / Torch7: "parity unpinned" for
the wiring).  Only tests/, smoke() and bench.py's checker legs may import this module.
"""
from __future__ import annotations

import numpy as np

from . import b2f_oracle as o

F32 = np.float32
FEAT = (3, 16, 32, 64, 96, 128, 192)        # featMaps, pwc.lua:89 (d = 16)
DEC = (128, 128, 96, 64, 32, 2)             # decoder widths, pwc.lua:76-85


class Opt:
    """The opts.lua fields createModelMulti reads (pwc.lua:100-113) with the reference's defaults (opts.lua:83-98)."""

    def __init__(self, **kw):
        self.pwc_ws = 9
        self.frames = 3
        self.levels = 7
        self.pwc_skip = 2
        self.flownet_factor = 20
        self.past_flow = False
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unsupported option %r (only the Ours-Hard / Ours-Soft family is built)" % k)
            setattr(self, k, v)

    @property
    def l_st(self):
        return max(self.pwc_skip + 1, 1)


# ---------------------------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------------------------

def conv3x3(x, w, b=None, stride=1, dtype=np.float64):
    """nn.SpatialConvolution(nIn, nOut, 3, 3, s, s, 1, 1):updateOutput.  x (B, Cin, H, W), w (Cout, Cin, 3, 3)."""
    x = np.asarray(x, dtype)
    w = np.asarray(w, dtype)
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    xp = np.zeros((B, Cin, H + 2, W + 2), dtype)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((B, Cout, Ho, Wo), dtype)
    for ky in range(3):
        for kx in range(3):
            win = xp[:, :, ky:ky + (Ho - 1) * stride + 1:stride, kx:kx + (Wo - 1) * stride + 1:stride]
            out += np.matmul(w[:, :, ky, kx][None], win.reshape(B, Cin, Ho * Wo)).reshape(B, Cout, Ho, Wo)
    if b is not None:
        out += np.asarray(b, dtype)[None, :, None, None]
    return out


def leaky_relu(x, slope=0.2):
    return np.where(x > 0, x, x * slope)


def avgpool2x2(x):
    B, C, H, W = x.shape
    Ho, Wo = H // 2, W // 2
    v = x[:, :, :2 * Ho, :2 * Wo].reshape(B, C, Ho, 2, Wo, 2)
    return v.sum(axis=(3, 5)) / 4.0


def upsample_bilinear2x(x, dtype=np.float64):
    """THNN SpatialUpSamplingBilinear: ratio (in - 1) / (out - 1), h1 = floor(r h2), h1p = h1 < H - 1, lambdas."""
    x = np.asarray(x, dtype)
    B, C, H, W = x.shape
    Ho, Wo = 2 * H, 2 * W

    def axis(n_in, n_out):
        r = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        src = r * np.arange(n_out)
        i0 = np.floor(src).astype(np.int64)
        i0 = np.minimum(i0, n_in - 1)
        p = (i0 < n_in - 1).astype(np.int64)
        l1 = src - i0
        return i0, p, l1

    h0, hp, lh = axis(H, Ho)
    w0, wp, lw = axis(W, Wo)
    a = x[:, :, h0][:, :, :, w0]
    b_ = x[:, :, h0][:, :, :, w0 + wp]
    c = x[:, :, h0 + hp][:, :, :, w0]
    d = x[:, :, h0 + hp][:, :, :, w0 + wp]
    lw_ = lw[None, None, None, :]
    lh_ = lh[None, None, :, None]
    return (1 - lh_) * ((1 - lw_) * a + lw_ * b_) + lh_ * ((1 - lw_) * c + lw_ * d)


def upsample_nearest(x, scale=2):
    return np.repeat(np.repeat(x, scale, axis=2), scale, axis=3)


def spatial_softmax(x):
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=1, keepdims=True)


# ---------------------------------------------------------------------------------------------------------
# parameters
# ---------------------------------------------------------------------------------------------------------

def conv_shapes(opt=None):
    """Ordered (name, Cout, Cin, stride) of every convolution of createModelMulti(opt) -- names are this repo's:
    feat.l<l>.<0|1> (the siamese convUnits, shared over the frames, pwc.lua:176-195), occ.l<l>.<i>, flow.l<l>.<i>,
    bflow.l<l>.<i> (decoders of level l, i = 0..5; pwc.lua:288-338)."""
    opt = opt or Opt()
    shapes = []
    for l in range(2, opt.levels + 1):
        shapes.append(("feat.l%d.0" % l, FEAT[l - 1], FEAT[l - 2], 2))
        shapes.append(("feat.l%d.1" % l, FEAT[l - 1], FEAT[l - 1], 1))
    nd = 2 * opt.pwc_ws ** 2
    for l in range(opt.levels, opt.l_st - 1, -1):
        n_occ = nd + FEAT[l - 1] + (2 if l != opt.levels else 0)
        n_flow = nd if l == opt.levels else nd + FEAT[l - 1] + 2
        for kind, n_in in (("occ", n_occ), ("flow", n_flow)) + ((("bflow", n_flow),) if opt.past_flow else ()):
            cin = n_in
            for i, cout in enumerate(DEC):
                shapes.append(("%s.l%d.%d" % (kind, l, i), cout, cin, 1))
                cin = cout
    return shapes


def init_params(opt=None, seed=2, scale=1.0):
    """Random weights with nn.SpatialConvolution:reset()'s distribution: uniform(-s, s), s = 1 / sqrt(9 nIn), for
    weight and bias (the pretrained .t7 files are not available offline).  `scale` > 1 widens it (tests use it to
    keep activations from fading through six layers)."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, cout, cin, _s in conv_shapes(opt):
        s = scale / np.sqrt(9.0 * cin)
        params[name + ".weight"] = rng.uniform(-s, s, (cout, cin, 3, 3)).astype(F32)
        params[name + ".bias"] = rng.uniform(-s, s, (cout,)).astype(F32)
    return params


def n_params(opt=None):
    return sum(co * ci * 9 + co for _n, co, ci, _s in conv_shapes(opt))


# ---------------------------------------------------------------------------------------------------------
# the network
# ---------------------------------------------------------------------------------------------------------

def conv_unit(params, name, x, dtype=np.float64):
    """convUnit(d_in, d_out, 2), pwc.lua:58-65."""
    y = leaky_relu(conv3x3(x, params[name + ".0.weight"], params[name + ".0.bias"], 2, dtype))
    return leaky_relu(conv3x3(y, params[name + ".1.weight"], params[name + ".1.bias"], 1, dtype))


def decoder(params, name, x, dtype=np.float64):
    """decoder(nChannels), pwc.lua:76-85: five conv + LeakyReLU(0.2), one plain conv to 2 channels."""
    for i in range(6):
        x = conv3x3(x, params["%s.%d.weight" % (name, i)], params["%s.%d.bias" % (name, i)], 1, dtype)
        if i < 5:
            x = leaky_relu(x)
    return x


def feature_pyramid(params, img, opt=None, dtype=np.float64):
    """cs[f][l], l = 1..levels, for one frame (B, 3, H, W); index l - 1 (pwc.lua:197-211, skip > 0: cs[f][1] = I)."""
    opt = opt or Opt()
    cs = [np.asarray(img, dtype)]
    for l in range(2, opt.levels + 1):
        cs.append(conv_unit(params, "feat.l%d" % l, cs[-1], dtype))
    return cs


def pwc_forward(params, x, opt=None, dtype=np.float64, taps=None):
    """createModelMulti(opt):forward(x), pwc.lua:139-492, frames = 3, two_frame = 0, pwc_sum_cvs = false,
    residual = 0, occ_input = 0, rescale_flow = 0, siamese = 1, skip > 0.

    x (B, 9, H, W), H and W multiples of 2^(levels-1).  Returns the output table, finest level first
    (:459-489): per level {skip_ufs, [skip_ubfs,] skip_occs, iws[1], iws[3]}.  `taps` (a dict) receives the
    intermediate tensors by name for layer-wise comparisons."""
    opt = opt or Opt()
    assert opt.frames == 3 and opt.pwc_skip > 0
    x = np.asarray(x, dtype)
    B, nine, H, W = x.shape
    assert nine == 9 and H % (1 << (opt.levels - 1)) == 0 and W % (1 << (opt.levels - 1)) == 0
    ref, l_st, levels, win = 2, opt.l_st, opt.levels, opt.pwc_ws
    Is = {f: x[:, 3 * (f - 1):3 * f] for f in (1, 2, 3)}                       # nn.Narrow(2, a, 3), :141-146
    ds = {}
    for f in (1, 3):                                                          # :149-158
        ds[f] = [Is[f]]
        for _l in range(2, levels - l_st + 2):
            ds[f].append(avgpool2x2(ds[f][-1]))
    cs = {f: feature_pyramid(params, Is[f], opt, dtype) for f in (1, 2, 3)}    # cs[f][l-1]
    ws = {1: {}, 3: {}}
    fs, bfs, ufs, ubfs, skip_ufs, skip_ubfs, skip_occs, iws = {}, {}, {}, {}, {}, {}, {}, {1: {}, 3: {}}
    tap = taps if taps is not None else {}
    for l in range(levels, l_st - 1, -1):
        inp = cs if l == levels else None
        refl = cs[ref][l - 1]
        fut = inp[3][l - 1] if inp else ws[3][l]
        past = inp[1][l - 1] if inp else ws[1][l]
        cv_f = o.costvol_forward([refl, fut], win, True, dtype)                # :252
        cv_b = o.costvol_forward([refl, past], win, False, dtype)              # :264
        cvs = np.concatenate([cv_f, cv_b], axis=1)                             # JoinTable(2), :267
        tap["cvs.l%d" % l] = cvs
        occ_in = [cvs, refl] + ([ufs[l + 1]] if l != levels else [])           # :288-302
        occ = spatial_softmax(decoder(params, "occ.l%d" % l, np.concatenate(occ_in, axis=1), dtype))   # :305
        tap["occ.l%d" % l] = occ
        so = upsample_nearest(occ, 2)                                          # uoccs, :308-310
        for _i in range(2, l_st):
            so = upsample_nearest(so, 2)                                       # :313-317
        skip_occs[l] = so
        if l == levels:                                                        # :322-328
            fs[l] = decoder(params, "flow.l%d" % l, cvs, dtype)
            if opt.past_flow:
                bfs[l] = decoder(params, "bflow.l%d" % l, cvs, dtype)
        else:                                                                  # :330-349
            fs[l] = decoder(params, "flow.l%d" % l, np.concatenate([cvs, refl, ufs[l + 1]], axis=1), dtype)
            if opt.past_flow:
                bfs[l] = decoder(params, "bflow.l%d" % l, np.concatenate([cvs, refl, ubfs[l + 1]], axis=1), dtype)
        tap["fs.l%d" % l] = fs[l]
        ufs[l] = upsample_bilinear2x(fs[l], dtype)                             # :358-362 (skip > 0)
        su = ufs[l]
        for _i in range(2, l_st):
            su = upsample_bilinear2x(su, dtype)                                # :374-389
        skip_ufs[l] = su
        if opt.past_flow:
            ubfs[l] = upsample_bilinear2x(bfs[l], dtype)
            sb = ubfs[l]
            for _i in range(2, l_st):
                sb = upsample_bilinear2x(sb, dtype)
            skip_ubfs[l] = sb
        for f in (1, 3):                                                       # :394-448
            if l > l_st:
                sc = opt.flownet_factor * (f - ref) / 2.0 ** (l - 2)           # :404
                ws[f][l - 1] = o.warping_unit_forward(_as32(cs[f][l - 2]), _as32(ufs[l]), sc, dtype)
                tap["ws%d.l%d" % (f, l - 1)] = ws[f][l - 1]
            tmp = skip_ubfs[l] if (opt.past_flow and f < ref) else skip_ufs[l]  # :426-437
            sc = opt.flownet_factor * (f - ref) / 2.0 ** (l - l_st)            # :443
            iws[f][l] = o.warping_unit_forward(_as32(ds[f][l - l_st]), _as32(tmp), sc, dtype)
    out = []
    for l in range(l_st, levels + 1):                                          # :459-489
        out.append(skip_ufs[l])
        if opt.past_flow:
            out.append(skip_ubfs[l])
        out.append(skip_occs[l])
        out.append(iws[1][l])
        out.append(iws[3][l])
    return out


def _as32(a):
    """The sampler's oracle takes fp32 inputs (it restates an fp32 kernel's address arithmetic: the product
    flow * scale and the floor of the coordinate are fp32 operations of the reference, BilinearSamplerBHWD.cu:6-20)."""
    return np.asarray(a, F32)


def flow_scales(opt=None):
    """model.flow_scale (pwc.lua:451-455, 493): coarse -> fine."""
    opt = opt or Opt()
    return [opt.flownet_factor / 2.0 ** (l - opt.l_st) for l in range(opt.levels, opt.l_st - 1, -1)]
