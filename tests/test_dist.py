"""CPU tests of the multi-GPU plumbing with world_size-2 gloo processes: sharding covers
every unit exactly once, and the sharded training reduction (per-rank criterion gradients on a batch
shard + sum all-reduce of the flattened gradient) equals the single-process result of the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from back2future_b200 import dist as bdist


def test_split_range_covers_everything_once():
    for n in (0, 1, 7, 8, 62, 63, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = bdist.split_range(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    sizes = [bdist.split_range(62, 8, r) for r in range(8)]
    assert [hi - lo for lo, hi in sizes] == [8, 8, 8, 8, 8, 8, 7, 7]
    with pytest.raises(ValueError):
        bdist.split_range(4, 2, 2)


def test_triplet_shard_frames_have_one_frame_halo():
    F, world = 64, 8
    all_trip = []
    for r in range(world):
        (tlo, thi), (flo, fhi) = bdist.triplet_shard(F, world, r)
        all_trip += list(range(tlo, thi))
        assert (flo, fhi) == (tlo, thi + 2)      # triplet t = frames t, t+1, t+2
    assert all_trip == list(range(F - 2))
    assert bdist.triplet_shard(2, 4, 1) == ((0, 0), (0, 0))  # no triplets at all
    (tlo, thi), frames = bdist.triplet_shard(3, 4, 0)
    assert (tlo, thi) == (0, 1) and frames == (0, 3)
    assert bdist.triplet_shard(3, 4, 3)[0] == (1, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import b2f_oracle as o
        rng = np.random.default_rng(2)
        B, h, w = 4, 6, 8
        flow = (rng.standard_normal((B, 2, h, w)) * 0.3).astype(np.float32)
        e = np.exp(rng.standard_normal((B, 2, h, w)))
        occ = (e / e.sum(1, keepdims=True)).astype(np.float32)
        w1, w2, tgt = (rng.uniform(-2, 2, (B, 3, h, w)).astype(np.float32) for _ in range(3))
        lo, hi = bdist.batch_shard(B, world, rank)
        crit = o.OBCriterionOracle(False, o.L1Penalty(), pwc_flow_scaling=20.0, size_average=False)
        sl = slice(lo, hi)
        loss = crit.forward(flow[sl], None, occ[sl], [w1[sl], w2[sl]], tgt[sl])
        g_occ, g_warp = crit.backward(flow[sl], None, occ[sl], [w1[sl], w2[sl]], tgt[sl])
        # a toy "network": the parameter gradient is a fixed linear map of the per-sample output
        # gradients, summed over the batch -- exactly the structure that makes DP gradients add.
        proj = np.random.default_rng(7).standard_normal((g_occ[0].size + 2 * g_warp[0][0].size, 16))
        per_sample = np.concatenate([g_occ.reshape(hi - lo, -1), g_warp[0].reshape(hi - lo, -1),
                                     g_warp[1].reshape(hi - lo, -1)], axis=1)
        flat = torch.from_numpy((per_sample @ proj).sum(axis=0))
        red = bdist.GradientAllReduce(flat, bucket_elems=5)
        assert red.world == world and len(red.buckets()) == 4
        red.start()
        red.wait()
        total_loss, = bdist.reduce_losses([loss])
        if rank == 0:
            np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
            np.save(os.path.join(out_dir, "loss.npy"), np.array([total_loss]))
            # single-process reference on the whole batch
            full = crit.forward(flow, None, occ, [w1, w2], tgt)
            go, gw = crit.backward(flow, None, occ, [w1, w2], tgt)
            ps = np.concatenate([go.reshape(B, -1), gw[0].reshape(B, -1), gw[1].reshape(B, -1)], axis=1)
            np.save(os.path.join(out_dir, "flat_ref.npy"), (ps @ proj).sum(axis=0))
            np.save(os.path.join(out_dir, "loss_ref.npy"), np.array([full]))
    finally:
        dist.destroy_process_group()


def test_sharded_training_reduction_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    flat, ref = np.load(tmp_path / "flat.npy"), np.load(tmp_path / "flat_ref.npy")
    assert np.allclose(flat, ref, rtol=1e-10, atol=1e-10)
    assert np.isclose(np.load(tmp_path / "loss.npy")[0], np.load(tmp_path / "loss_ref.npy")[0], rtol=1e-12)


def test_world_size_one_is_a_noop():
    flat = torch.arange(10, dtype=torch.float32)
    red = bdist.GradientAllReduce(flat)
    red.start()
    red.wait()
    assert torch.equal(flat, torch.arange(10, dtype=torch.float32))
    assert bdist.reduce_losses([1.5, 2.5]) == [1.5, 2.5]
    assert bdist.NPARAMS_HARD == 7193316 and bdist.NPARAMS_SOFT == 10168302
