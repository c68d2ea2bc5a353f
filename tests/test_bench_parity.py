"""Parity gate ON THE THING THAT IS BENCHMARKED: build bench.py's own Workload, replay the captured CUDA graph that
`value` times, and compare all 56 ops (cost volume L3-L7 forward + both gradients, both directions; 8 feature warps;
10 image warps), every batch item, backward included, with the float64 checker at 1e-4 -- at B = 8 / 1024x448
(BASELINE configs[1]) and at the B = 1 shapes of configs[0] (1216x320) and configs[4] (1024x384)."""
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,full_hw", [(8, None), (1, (384, 1024)), (1, (320, 1216)), (2, (192, 320))])
@pytest.mark.parametrize("launch", ["graph", "eager"])
def test_every_op_of_the_benchmarked_step_matches_the_oracle(cuda_lib, B, full_hw, launch):
    import torch
    import bench
    from oracle import workload_check

    if launch == "eager" and B == 8:
        pytest.skip("the eager order issues the same 56 calls; checked at the smaller shapes")
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    wl = bench.Workload(torch, cuda_lib, dev, B=B, full_hw=full_hw)
    wl.bind(torch.cuda.current_stream().cuda_stream)
    wl.step()                         # warm-up: first call of every kernel sets its attributes
    torch.cuda.synchronize()
    for _, t in wl.outputs:           # poison: a kernel that writes nothing cannot pass
        t.fill_(float("nan"))
    if launch == "graph":
        streams = [torch.cuda.Stream(device=dev) for _ in range(6)]
        graph = wl.capture(streams)
        graph.replay()
    else:
        wl.step()
    torch.cuda.synchronize()
    r = workload_check.check_workload(wl, verbose=sys.stderr.write)
    assert r["checked"] == 10 * 3 + 18 * 3 and r["ops"] == 56
    assert not r["failed"], r
    assert r["max_rel_err"] < 1e-4
    del wl
    torch.cuda.empty_cache()


def test_wrong_results_are_caught(cuda_lib):
    """The gate must fail when a kernel is in one of its measurement modes (b2f_debug_costvol_path 8: the tiled
    kernels without their arithmetic -- 'produce WRONG results', include/b2f.h)."""
    import torch
    import bench
    from oracle import workload_check

    dev = torch.device("cuda:0")
    wl = bench.Workload(torch, cuda_lib, dev, B=2, full_hw=(192, 512))
    wl.bind(torch.cuda.current_stream().cuda_stream)
    prev = cuda_lib.b2f_debug_costvol_path(8)
    try:
        wl.step()
        torch.cuda.synchronize()
    finally:
        cuda_lib.b2f_debug_costvol_path(prev)
    r = workload_check.check_workload(wl)
    assert r["failed"] and all(n.startswith("costvol") for n in r["failed"])


def test_e2e_leg_returns_the_modules_results(cuda_lib):
    """bench.py's end-to-end leg (one pinned input arena up, module calls on views of the device arena, one output arena
    down; two sets of device arenas): what arrives in the host output arena after several pipelined steps is what the
    float64 checker computes from the host input arena -- first cost-volume item, first feature warp, last image warp."""
    import numpy as np
    import torch
    import bench
    from oracle import check64 as c64

    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    ee = bench.E2E(torch, dev, B=1)
    ee.h_out.fill_(float("nan"))
    for _ in range(3):                # both arena slots, and a slot reused
        ee.step()
    ee.drain()
    torch.cuda.synchronize()
    hin, hout = ee.h_in.numpy(), ee.h_out.numpy()

    def numel(t):
        return int(np.prod(t.shape))

    oi = oo = 0
    checked = 0
    for idx, (kind, _mod, vin, vout) in enumerate(ee.items):
        ins, outs = [], []
        for v in vin[0]:
            ins.append(hin[oi:oi + numel(v)].reshape(tuple(v.shape)))
            oi += numel(v)
        for v in vout[0]:
            outs.append(hout[oo:oo + numel(v)].reshape(tuple(v.shape)))
            oo += numel(v)
        assert all(np.isfinite(o_).all() for o_ in outs), (idx, kind)
        if kind == "cv" and idx == 0:
            ref, past, fut, gj = ins
            want = np.concatenate([c64.costvol_forward([ref, fut], 9, True), c64.costvol_forward([ref, past], 9, False)], 1)
            assert c64.rel_err(outs[0], want) < 1e-4
            g = c64.costvol_backward([ref, fut], gj[:, :81], 9, True)
            assert c64.rel_err(outs[1], g[0]) < 1e-4 and c64.rel_err(outs[2], g[1]) < 1e-4
            g = c64.costvol_backward([ref, past], gj[:, 81:], 9, False)
            assert c64.rel_err(outs[3], g[0]) < 1e-4 and c64.rel_err(outs[4], g[1]) < 1e-4
            checked += 1
        if kind == "warp" and (idx == 5 or idx == len(ee.items) - 1):
            img, grid, go = ins
            assert c64.rel_err(outs[0], c64.warp_forward(img, grid)) < 1e-4
            gi, gg = c64.warp_backward(img, grid, go)
            assert c64.rel_err(outs[1], gi) < 1e-4 and c64.rel_err(outs[2], gg) < 1e-4
            checked += 1
    assert checked == 3 and oi == ee.h_in.numel() and oo == ee.h_out.numel()
