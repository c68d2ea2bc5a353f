"""Parity gate of the benchmarked step -- TEST INFRASTRUCTURE, NOT PRODUCT.

`check_workload(wl)` compares EVERY op of a `bench.Workload` (cost volume forward + both gradients for both
directions at every level, every feature warp, every image warp: 56 ops at the default pyramid) on the buffers the
timed graph has just written, with the float64 checker (oracle/check64.py, pinned to the numpy oracle by
tests/test_oracle.py), every batch item, at the workload's full size.  Used by tests/test_bench_parity.py and by
bench.py's `parity` block (after its timed region; the oracle is the checker there, never the thing measured).
"""
from __future__ import annotations

import numpy as np

from . import check64 as c64

TOL = 1e-4          # north_star: 1e-4 relative for fp32 values and gradients
TOL_ATOMIC = 1e-4   # image gradient of the sampler: float atomics, same bar


def check_workload(wl, tol=TOL, verbose=None):
    """Returns {"checked": n_ops, "max_rel_err": e, "worst": name, "failed": [names], "tol": tol}."""
    torch = wl.torch
    torch.cuda.synchronize()
    inp = {n: t for n, t in wl.inputs}
    out = {n: t for n, t in wl.outputs}
    host = lambda t: t.detach().cpu().numpy()
    res = []

    def rec(name, kernel, ref):
        e = c64.rel_err(kernel, ref)
        res.append((name, e))
        if verbose:
            verbose("%-40s %.3e\n" % (name, e))

    levels = sorted({int(n[2:].split(".")[0]) for n in inp if n.startswith("cv")}, reverse=True)
    for l in levels:
        ref, past, fut = (host(inp["cv%d.%s" % (l, k)]) for k in ("ref", "past", "fut"))
        gj = host(inp["cv%d.gradJoined" % l])
        joined = host(out["cv%d.joined" % l])
        for d, (frame, fwd, half, nm) in enumerate(((fut, True, 0, "fut"), (past, False, 1, "past"))):
            sl = slice(81 * half, 81 * (half + 1))
            rec("costvol_fwd L%d %s" % (l, nm), joined[:, sl], c64.costvol_forward([ref, frame], 9, fwd))
            g = c64.costvol_backward([ref, frame], gj[:, sl], 9, fwd)
            rec("costvol_bwd L%d %s gradRef" % (l, nm), host(out["cv%d.grad%d" % (l, 2 * d)]), g[0])
            rec("costvol_bwd L%d %s gradFrame" % (l, nm), host(out["cv%d.grad%d" % (l, 2 * d + 1)]), g[1])
    for n in [n for n, _ in wl.inputs if n.endswith(".img")]:
        tag = n[:-4]
        img, grid, go = host(inp[tag + ".img"]), host(inp[tag + ".grid"]), host(inp[tag + ".gradOut"])
        rec(tag + " fwd", host(out[tag + ".out"]), c64.warp_forward(img, grid))
        gi, gg = c64.warp_backward(img, grid, go)
        rec(tag + " bwd gradImg", host(out[tag + ".gradImg"]), gi)
        rec(tag + " bwd gradGrid", host(out[tag + ".gradGrid"]), gg)
    worst = max(res, key=lambda r: r[1])
    return {"checked": len(res), "ops": len(wl._mk), "max_rel_err": float("%.3e" % worst[1]), "worst": worst[0],
            "failed": [n for n, e in res if not e < tol], "tol": tol,
            "oracle": "oracle/c/b2f_check64.c (float64 closed forms, pinned to oracle/b2f_oracle.py)"}
