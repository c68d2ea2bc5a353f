"""Data-parallel training step across ranks (torchrun): the one-graph step (bucket all-reduces captured into the graph)
against the eager step (same calls, collectives issued from the host), same initial parameters, two steps each.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_train_multi.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from back2future_b200 import pwc, train, comm as bcomm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
cm = bcomm.Communicator.from_env()
B, H, W = 2, 128, 192
g = torch.Generator().manual_seed(100 + rank)
x = torch.empty(B, 9, H, W).uniform_(-2.1, 2.6, generator=g)
res = {}
for mode in ("graph", "eager"):
    torch.manual_seed(7)
    net = pwc.PWCNet(pwc.Opt(past_flow=False), device=dev, image_warps=True, seed=3)
    tr = train.Trainer(net, train.TrainOpt.hard(), comm=cm)
    for _ in range(2):
        losses = tr.train_batch(x, graph=(mode == "graph"))
    res[mode] = (net.flat_params.clone(), net.flat_grads.clone(), losses["err"])
    del tr, net
pg, gg, lg = res["graph"]
pe, ge, le = res["eager"]
rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
# every rank must hold the same reduced gradient and parameters
ref = [torch.empty_like(gg) for _ in range(world)]
dist.all_gather(ref, gg)
same = max(rel(r, ref[0]) for r in ref)
print("rank %d: loss graph %.6f eager %.6f | grads graph vs eager %.2e | params %.2e | grads across ranks %.2e"
      % (rank, lg, le, rel(gg, ge), rel(pg, pe), same), flush=True)
# Parameters after two Adam steps: the update of an element is lr * m / (sqrt(v) + eps), i.e. ~ +-lr whatever the size
# of its gradient, so an element whose gradient is at the level of the run-to-run rounding noise (atomic order: 1e-7
# of the largest gradient) can move by 2 lr the other way.  Bound: no element differs by more than 2 lr per step, and
# all but a vanishing fraction agree to 1e-4.
lr = train.TrainOpt.hard().LR
dp = (pg - pe).abs()
frac = (dp > 1e-4 * pe.abs().max()).float().mean().item()
print("rank %d: params max abs diff %.2e (bound %.2e), fraction beyond 1e-4: %.2e" % (rank, dp.max().item(), 4.04 * lr, frac), flush=True)
ok = rel(gg, ge) < 1e-4 and dp.max().item() <= 4.04 * lr and frac < 1e-3 and same == 0.0
dist.barrier()
cm.destroy()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
