"""Multi-GPU plumbing for the hot path: one process per GPU, `torch.distributed` (NCCL over NVLink 5 /
NVSwitch on the B200 box, gloo in the CPU tests).

What the reference does (util.lua:27-48, train.lua:480-496): `nn.DataParallelTable(1, true, true)` --
split the batch on dim 1 across GPUs inside ONE process, reduce the flattened gradient to GPU 1 with
nccl.torch, broadcast the parameters back, and run every criterion on GPU 1 for the gathered batch.

What this module does instead:
  * inference shards frame triplets across ranks -- contiguous ranges with a one-frame halo, no
    inter-GPU traffic at all;
  * training shards the batch; every rank runs the cost volumes, warps AND its own criterions on
    its shard (all of them are per-sample except the Q9 aliasing inside one criterion call, which
    stays inside a shard); `sizeAverage=false` losses are sums, so gradients add across ranks with
    no rescale: one all-reduce(sum) of the flattened fp32 gradient (7.19 M floats for Ours-Hard,
    10.17 M with past_flow) per step, issued on a side stream so it overlaps the tail of backward;
  * loss scalars are summed with one tiny all-reduce for logging.
No collective is invented where the path has none.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

# parameter counts of the reference architectures (derived from models/pwc.lua)
NPARAMS_HARD = 7_193_316
NPARAMS_SOFT = 10_168_302


def split_range(n, world, rank):
    """Contiguous [lo, hi) share of n items: the first n % world ranks get one extra item
    (64 frames -> 62 triplets -> 8,8,8,8,8,8,7,7 on 8 ranks)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def triplet_shard(num_frames, world, rank):
    """Inference sharding of a frame sequence.  Triplet t (0-based) uses frames (t, t+1, t+2) with
    reference frame t+1; a sequence of F frames has F-2 triplets.  Returns
    ((t_lo, t_hi), (frame_lo, frame_hi)): the rank's triplets and the frames it must load
    (its triplets' frames = a one-frame halo on each side of the reference frames)."""
    ntrip = max(0, num_frames - 2)
    lo, hi = split_range(ntrip, world, rank)
    if hi <= lo:
        return (lo, lo), (0, 0)
    return (lo, hi), (lo, hi + 2)


def batch_shard(batch, world, rank):
    """Training sharding of a global batch on dim 0 (what DataParallelTable(1, ...) does)."""
    return split_range(batch, world, rank)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class GradientAllReduce:
    """Sum-all-reduce of the flattened gradient buffer (the reference's `flattenParams=true,
    usenccl=true` reduction).  On CUDA the collective runs on its own stream: `start()` is called
    as soon as backward has produced the gradient (or a bucket of it), `wait()` before the
    optimizer step; with world size 1 both are no-ops."""

    def __init__(self, flat_grad: torch.Tensor, bucket_elems: int | None = None):
        if flat_grad.dim() != 1 or not flat_grad.is_contiguous():
            raise ValueError("flat_grad must be a contiguous 1-D tensor")
        self.flat = flat_grad
        self.rank, self.world = world_info()
        n = flat_grad.numel()
        # NVSwitch gives every peer full bandwidth and NCCL reduces in-switch (NVLS): buckets are
        # sized for launch latency / overlap only, not for link count.  Default: one message.
        self.bucket = n if not bucket_elems else max(1, min(int(bucket_elems), n))
        self._handles = []
        self._stream = torch.cuda.Stream(device=flat_grad.device) if flat_grad.is_cuda else None
        self._ready = None

    def buckets(self):
        n = self.flat.numel()
        return [(lo, min(n, lo + self.bucket)) for lo in range(0, n, self.bucket)]

    def start(self, lo=0, hi=None):
        """Launch the reduction of flat[lo:hi] (default: everything)."""
        if self.world == 1:
            return
        hi = self.flat.numel() if hi is None else hi
        view = self.flat[lo:hi]
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self._stream):
                for blo, bhi in self._chunks(lo, hi):
                    self._handles.append(dist.all_reduce(self.flat[blo:bhi], op=dist.ReduceOp.SUM, async_op=True))
        else:
            for blo, bhi in self._chunks(lo, hi):
                self._handles.append(dist.all_reduce(self.flat[blo:bhi], op=dist.ReduceOp.SUM, async_op=True))
        del view

    def _chunks(self, lo, hi):
        return [(a, min(hi, a + self.bucket)) for a in range(lo, hi, self.bucket)]

    def wait(self):
        for h in self._handles:
            h.wait()
        self._handles = []
        if self._stream is not None and self.world > 1:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._stream)


def reduce_losses(values, device=None):
    """Sum a list of per-rank loss scalars (python floats) across ranks; returns python floats.
    Losses are sums (`sizeAverage=false`, model.lua:249-258), so the global loss is the plain sum."""
    rank, world = world_info()
    if world == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]
