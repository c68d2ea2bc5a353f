"""`back2future.init(name)` / `computeFlow(im1, im2, im3)` of the reference (back2future.lua:47-129, README.md:49-60)
over the B200 path, plus the sequence driver the reference leaves to the user (a loop over frame triplets).

    b2f = Back2Future.init("Ours-Hard")                # random-init unless models/RoamingImages_H.t7 exists
    flow, fwd_occ, bwd_occ = b2f.computeFlow(im1, im2, im3)      # (3, H, W) float arrays in [0, 1]

Host side (exactly where the reference has it, back2future.lua:48-72 and :76-92 run on host tensors): channel
concatenation, ColorNormalize, `image.scale` to a multiple of 64, and after the network the 'simple' resize, the flow
rescale and the double-precision occlusion threshold.  Device side: one PWCNet forward (a CUDA-graph replay).

Quirk Q12: the reference reads the occlusion map from `est[3]`, which is correct only for the Soft
(past_flow) models; for Ours-Hard `est[3]` is warped frame 1.  `occ_index="as_written"` reproduces that literally
(channels 1 and 2 of whatever est[3] is), the default `"occlusion"` takes the occlusion entry of the output unit.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import imageio, pwc, t7
from .dist import triplet_shard

MODEL_FILES = {                                   # back2future.lua:100-110
    "Ours-Hard": ("RoamingImages_H.t7", False),
    "Ours-Soft-ft-KITTI": ("RoamingImages_H_KITTI_S.t7", True),
    "Ours-Soft-ft-Sintel": ("RoamingImages_H_Sintel_S.t7", True),
}


class Back2Future:
    def __init__(self, net, occ_index="occlusion"):
        if occ_index not in ("occlusion", "as_written"):
            raise ValueError("occ_index must be 'occlusion' or 'as_written'")
        self.model = net
        self.occ_index = occ_index
        self.channels = 9                          # 3 * #nn.Narrow, back2future.lua:120-126
        self._pinned = {}

    # ---- back2future.init ---------------------------------------------------------------------------------
    @classmethod
    def init(cls, opt="Ours-Soft-ft-KITTI", model_dir="models", device="cuda:0", seed=2, occ_index="occlusion",
             image_warps=False, tensor_cores=True):
        """back2future.lua:97-129.  The checkpoint is read with `t7.load` when the file is there; the published weights
        are not available offline, so otherwise the architecture is built with nn.SpatialConvolution:reset()'s random
        initialisation (stated in every benchmark line as `data: synthetic`).  `image_warps=False`: computeFlow never
        reads the warped frames (est[4..5]), so the plan leaves those launches out."""
        if opt not in MODEL_FILES:
            raise ValueError("unknown model %r (expected one of %s)" % (opt, ", ".join(MODEL_FILES)))
        fname, past_flow = MODEL_FILES[opt]
        path = os.path.join(model_dir, fname)
        params = None
        if os.path.exists(path):
            params, pf = t7.import_model(t7.load(path))
            past_flow = pf
        net = pwc.PWCNet(pwc.Opt(past_flow=past_flow), params, device=device, seed=seed, image_warps=image_warps,
                         tensor_cores=tensor_cores)
        return cls(net, occ_index)

    # ---- computeFlow ----------------------------------------------------------------------------------------
    def prepare(self, im1, im2, im3):
        """back2future.lua:48-72: cat, ColorNormalize, image.scale to the network size.  Returns ((9, fh, fw) float32,
        (width, height) of the input)."""
        ims = [np.asarray(i, np.float32) for i in (im1, im2, im3)]
        for i in ims:
            if i.ndim != 3 or i.shape[0] != 3 or i.shape != ims[0].shape:
                raise ValueError("computeFlow: expected three (3, H, W) images of one size")
        imgs = imageio.color_normalize(np.concatenate(ims, axis=0))
        height, width = imgs.shape[1], imgs.shape[2]
        fw, fh = imageio.fine_size(width, height)
        if fw <= 0 or fh <= 0:
            raise ValueError("computeFlow: images smaller than 64 x 64 cannot be rounded to the network size")
        if (fw, fh) != (width, height):
            imgs = imageio.scale(imgs, fw, fh)
        return np.ascontiguousarray(imgs), (width, height)

    def _pinned_buf(self, shape):
        if shape not in self._pinned:
            self._pinned[shape] = torch.empty(shape, dtype=torch.float32).pin_memory()
        return self._pinned[shape]

    def network(self, imgs9):
        """`model:forward(imgs:resize(1, 9, fh, fw):cuda())` (back2future.lua:73-74): host (9, fh, fw) -> output table."""
        buf = self._pinned_buf((1,) + imgs9.shape)
        buf[0].copy_(torch.from_numpy(imgs9))
        return self.model.forward(buf)

    def finish(self, est, size):
        """back2future.lua:76-92: flow from est[1], occlusions from est[3] (see Q12), resized with 'simple' and
        rescaled; the threshold is evaluated in double."""
        width, height = size
        flow_d = est[0][0]
        if self.occ_index == "as_written":
            if not self.model.image_warps:
                raise RuntimeError("occ_index='as_written' reads est[3], which for Ours-Hard is warped frame 1: build "
                                   "the model with image_warps=True")
            occ_d = est[2][0][:2]                   # est[3]: the occlusion map only for past_flow models
        else:
            occ_d = est[2 if self.model.past_flow else 1][0]
        flow = flow_d.to("cpu", torch.float64).numpy()
        occ = occ_d.to("cpu", torch.float64).numpy()
        sc_h, sc_w = height / flow.shape[1], width / flow.shape[2]
        flow = imageio.scale(flow, width, height, "simple")
        flow[1] *= sc_h
        flow[0] *= sc_w
        fwd, bwd = imageio.occlusion_masks(occ)
        return flow, imageio.scale(fwd, width, height, "simple"), imageio.scale(bwd, width, height, "simple")

    def computeFlow(self, im1, im2, im3):
        imgs, size = self.prepare(im1, im2, im3)
        est = self.network(imgs)
        return self.finish(est, size)

    __call__ = computeFlow

    # ---- a frame sequence: triplets (t, t+1, t+2), sharded across ranks ---------------------------
    def compute_sequence(self, frames, rank=0, world=1):
        """Flow + occlusions for every triplet of this rank's contiguous share of `frames` (a list of (3, H, W) images
        or a callable index -> image).  Every frame is normalised / rescaled ONCE and uploaded ONCE (it appears in up
        to three triplets); the network input of a triplet is assembled on the device from the three resident frames,
        and the upload of frame t+3 is in flight while triplet t runs.  Returns a list of (t, flow, fwd_occ, bwd_occ)."""
        n = len(frames)
        get = frames.__getitem__
        (t_lo, t_hi), (f_lo, f_hi) = triplet_shard(n, world, rank)
        if t_hi <= t_lo:
            return []
        dev = self.model.device
        copy_stream = torch.cuda.Stream(dev)
        first = np.asarray(get(f_lo), np.float32)
        height, width = first.shape[1], first.shape[2]
        fw, fh = imageio.fine_size(width, height)

        def prep(i):
            im = imageio.color_normalize(np.asarray(get(i), np.float32))
            if (fw, fh) != (width, height):
                im = imageio.scale(im, fw, fh)
            return im

        ring = [torch.empty((3, fh, fw), device=dev) for _ in range(4)]
        pins = [torch.empty((3, fh, fw), dtype=torch.float32).pin_memory() for _ in range(4)]
        ready = [None] * 4

        def upload(i):
            s = (i - f_lo) % 4
            pins[s].copy_(torch.from_numpy(prep(i)))
            with torch.cuda.stream(copy_stream):
                ring[s].copy_(pins[s], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            ready[s] = ev

        for i in range(f_lo, min(f_lo + 3, f_hi)):
            upload(i)
        plan = self.model.plan(1, fh, fw)
        results = []
        cur = torch.cuda.current_stream(dev)
        for t in range(t_lo, t_hi):
            for k in range(3):
                s = (t + k - f_lo) % 4
                cur.wait_event(ready[s])
                plan.x[0, 3 * k:3 * k + 3].copy_(ring[s], non_blocking=True)       # D2D cudaMemcpyAsync
            copied = torch.cuda.Event()
            copied.record(cur)
            self.model.run(plan)                   # asynchronous: the host prepares the next frame meanwhile
            if t + 3 < f_hi:
                copy_stream.wait_event(copied)     # slot of frame t + 3 == slot of frame t - 1: last read by triplet t - 1
                upload(t + 3)
            results.append((t,) + self.finish(plan.output, (width, height)))
        return results


def init(opt="Ours-Soft-ft-KITTI", **kw):
    """`computeFlow = back2future.init(opt)` -- returns the callable, like the reference."""
    return Back2Future.init(opt, **kw)
