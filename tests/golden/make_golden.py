#!/usr/bin/env python
"""Generates tests/golden/hotpath_golden.npz  --  run in the BUILD container only.

    python tests/golden/make_golden.py [/root/reference]

What these vectors are (and are not).  The reference is Lua/Torch7 and cannot be executed here
(SURVEY 8c), so the expected values are produced by the float64 numpy oracle (oracle/b2f_oracle.py),
not by the reference itself: parity stays "unpinned".  The fixture pins the *oracle* (and through it
both product and C restatement) against silent change, and it carries real-image inputs -- crops of the
reference's own sample frames samples/frame_0009..0011.png, the inputs of its README inference example
(README.md:49-71, back2future.lua:47-95) -- to the GPU box, where /root/reference does not exist.

Inputs are built the way the network builds them:
  * frames  : ColorNormalize(mean/std of back2future.lua:33-36) of an average-pooled crop, (B,3,h,w);
  * features: a fixed seeded 3x3 "conv" (C=8) of each frame, a stand-in for the siamese convUnit
              (pwc.lua:58-65) with the same BDHW layout;
  * flows   : smooth fields (low-frequency sinusoids + seeded noise) in network units, so that
              flow*flow_scale leaves the image near the borders (the out-of-image mask and the sampler
              clamp are exercised);
  * occ     : channel softmax of seeded normals (pwc.lua:308);
  * warped frames: the oracle's own warp of the past / future frame with -flow / +flow.
Everything is stored as float32 (inputs exactly, expected outputs rounded from float64: 6e-8 relative,
far below the 1e-4 bar).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import b2f_oracle as o  # noqa: E402

MEAN = np.array([0.485, 0.456, 0.406], np.float32)
STD = np.array([0.229, 0.224, 0.225], np.float32)


def load_frames(ref_root, y0, x0, h, w, pool):
    from PIL import Image
    out = []
    for n in (9, 10, 11):
        im = np.asarray(Image.open(os.path.join(ref_root, "samples", "frame_%04d.png" % n)).convert("RGB"),
                        dtype=np.float32) / np.float32(255)
        crop = im[y0:y0 + h * pool, x0:x0 + w * pool]
        crop = crop.reshape(h, pool, w, pool, 3).mean(axis=(1, 3), dtype=np.float32)
        out.append(((crop - MEAN) / STD).transpose(2, 0, 1))
    return np.stack(out)   # (3 frames, 3, h, w)


def features(frames_b3hw, Cn, rng):
    """(B,3,h,w) -> (B,Cn,h,w): fixed seeded 3x3 filter bank + LeakyReLU(0.2), zero padding."""
    B, _, h, w = frames_b3hw.shape
    k = (rng.standard_normal((Cn, 3, 3, 3)) * 0.4).astype(np.float32)
    pad = np.pad(frames_b3hw, ((0, 0), (0, 0), (1, 1), (1, 1)))
    out = np.zeros((B, Cn, h, w), np.float32)
    for dy in range(3):
        for dx in range(3):
            out += np.einsum("oc,bchw->bohw", k[:, :, dy, dx], pad[:, :, dy:dy + h, dx:dx + w]).astype(np.float32)
    return np.where(out > 0, out, np.float32(0.2) * out).astype(np.float32)


def smooth_flow(B, h, w, amp, rng):
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    f = np.empty((B, 2, h, w), np.float32)
    for b in range(B):
        ph = rng.uniform(0, 2 * np.pi, 4)
        f[b, 0] = amp * (np.sin(2 * np.pi * xx / w + ph[0]) + 0.5 * np.cos(2 * np.pi * yy / h + ph[1]))
        f[b, 1] = amp * (0.7 * np.cos(2 * np.pi * xx / w + ph[2]) - np.sin(2 * np.pi * yy / h + ph[3]))
    return (f + rng.standard_normal(f.shape) * 0.1 * amp).astype(np.float32)


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    rng = np.random.default_rng(2)   # opts.lua:27 -manualSeed 2
    G = {}

    # two crops of the 1242x375 sample frames: (y0, x0) chosen on textured regions (cars / road edge)
    h, w = 20, 40
    crops = [load_frames(ref_root, 150, 300, h, w, 4), load_frames(ref_root, 120, 700, h, w, 4)]
    past, ref, fut = (np.stack([c[i] for c in crops]).astype(np.float32) for i in range(3))   # (2,3,h,w)
    G["frame_past"], G["frame_ref"], G["frame_fut"] = past, ref, fut

    # ---- cost volume (CostVolMulti.lua:49-181), C=8, both directions, gradOut a narrow of 162 channels
    Cn = 8
    fb = features(np.concatenate([past, ref, fut]), Cn, rng)
    f_past, f_ref, f_fut = fb[0:2], fb[2:4], fb[4:6]
    G["feat_fut_full"] = f_fut                                    # 20x40, used by the feature warp below
    # the cost-volume case runs on a 12x24 sub-crop (B=2) to keep the fixture small
    cpast, cref, cfut = (np.ascontiguousarray(a[:, :, 4:16, 8:32]) for a in (f_past, f_ref, f_fut))
    G["feat_past"], G["feat_ref"], G["feat_fut"] = cpast, cref, cfut
    go = rng.standard_normal((2, 162, 12, 24)).astype(np.float32)
    G["cv_gradout_joined"] = go
    G["cv_fwd_out"] = o.costvol_forward([cref, cfut], 9, True)
    G["cv_bwd_out"] = o.costvol_forward([cref, cpast], 9, False)
    g = o.costvol_backward([cref, cfut], go[:, :81], 9, True)
    G["cv_fwd_gradref"], G["cv_fwd_gradframe"] = g
    g = o.costvol_backward([cref, cpast], go[:, 81:], 9, False)
    G["cv_bwd_gradref"], G["cv_bwd_gradframe"] = g

    # ---- sampler (BilinearSamplerBHWD.cu:41-115, 161-307): image warp (C=3) and feature warp (C=8)
    flow = smooth_flow(2, h, w, 0.12, rng)            # network units; x20 = up to ~ +-4.5 px
    bflow = smooth_flow(2, h, w, 0.12, rng)
    G["flow"], G["bflow"] = flow, bflow
    scale = np.float32(20.0)
    grid_f = np.ascontiguousarray((flow * scale).transpose(0, 2, 3, 1))         # MulConstant(20*(+1)), BHW2 (x, y)
    grid_p = np.ascontiguousarray((flow * -scale).transpose(0, 2, 3, 1))        # MulConstant(20*(-1)) (pwc.lua:443)
    G["grid_fut"], G["grid_past"] = grid_f, grid_p
    img_f = np.ascontiguousarray(fut.transpose(0, 2, 3, 1))
    img_p = np.ascontiguousarray(past.transpose(0, 2, 3, 1))
    ft_f = np.ascontiguousarray(f_fut.transpose(0, 2, 3, 1))
    G["warp_img_fut"] = o.warp_forward(img_f, grid_f)
    G["warp_img_past"] = o.warp_forward(img_p, grid_p)
    G["warp_feat_fut"] = o.warp_forward(ft_f, grid_f)
    go3 = rng.standard_normal(img_f.shape).astype(np.float32)
    go8 = rng.standard_normal(ft_f.shape).astype(np.float32)
    G["warp_gradout3"], G["warp_gradout8"] = go3, go8
    G["warp_img_fut_gradimg"], G["warp_img_fut_gradgrid"] = o.warp_backward(img_f, grid_f, go3)
    G["warp_feat_fut_gradimg"], G["warp_feat_fut_gradgrid"] = o.warp_backward(ft_f, grid_f, go8)

    # ---- criterions on the warped real frames (train.lua:416-475 call order)
    w_past = np.ascontiguousarray(np.asarray(G["warp_img_past"], np.float32).transpose(0, 3, 1, 2))
    w_fut = np.ascontiguousarray(np.asarray(G["warp_img_fut"], np.float32).transpose(0, 3, 1, 2))
    e = np.exp(rng.standard_normal((2, 2, h, w)))
    occ = (e / e.sum(1, keepdims=True)).astype(np.float32)
    G["occ"], G["crit_warp_past"], G["crit_warp_fut"] = occ, w_past, w_fut
    for name, gt, past_flow, alpha in (("obcc", False, False, 1.0), ("obgcc", True, True, 0.0)):
        oc = o.OBCriterionOracle(gt, o.L1Penalty(), past_flow=past_flow, pwc_flow_scaling=20.0, size_average=False,
                                 alpha=alpha)
        bf = bflow if past_flow else None
        G[name + "_loss"] = np.float64(oc.forward(flow, bf, occ, [w_past, w_fut], ref))
        ro, rw = oc.backward(flow, bf, occ, [w_past, w_fut], ref)
        G[name + "_gradocc"], G[name + "_gradwarp_past"], G[name + "_gradwarp_fut"] = ro, rw[0], rw[1]
    for name, order, inp, pen in (("smooth1_flow", 1, flow, 1), ("smooth2_flow", 2, flow, 1), ("smooth1_occ", 1, occ, 0)):
        oc = o.SmoothnessOracle(order, o.make_penalty(pen), size_average=False, alias=True)
        G[name + "_loss"] = np.float64(oc.forward(inp, ref))
        G[name + "_grad"] = oc.backward(inp, ref)
    oc = o.SmoothnessOracle(1, o.L1Penalty(), size_average=False, alias=False)   # the evidently intended weights
    G["smooth1_flow_intended_loss"] = np.float64(oc.forward(flow, ref))
    G["smooth1_flow_intended_grad"] = oc.backward(flow, ref)
    G["constvel_loss"] = np.float64(o.constvel_forward(flow, bflow, True))
    G["constvel_gradf"], G["constvel_gradb"] = o.constvel_backward(flow, bflow, True)
    G["occprior_loss"] = np.float64(o.occprior_forward(occ, False))
    G["occprior_grad"] = o.occprior_backward(occ, False)
    G["mask_fut"] = o.out_of_image_mask(flow, 1, 20.0)
    G["mask_past_bflow"] = o.out_of_image_mask(bflow, -1, 20.0)

    out = {}
    for k, v in G.items():
        v = np.asarray(v)
        if v.dtype == np.float64 and v.ndim > 0:
            v = v.astype(np.float32)
        out[k] = v
    path = os.path.join(HERE, "hotpath_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
    print("mask_fut out-of-image pixels:", int((~G["mask_fut"]).sum()), "of", G["mask_fut"].size)


if __name__ == "__main__":
    main()
