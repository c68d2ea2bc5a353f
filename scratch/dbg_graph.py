import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from back2future_b200 import pwc
net = pwc.PWCNet(pwc.Opt())
B,H,W = 2,64,64
x = torch.randn(B,9,H,W,device="cuda")
p = net.plan(B,H,W)
def snap():
    torch.cuda.synchronize()
    d = {}
    for l in p.J: d["J%d"%l] = p.J[l].clone()
    for l in p.fs: d["fs%d"%l] = p.fs[l][0].clone()
    for l in p.occ: d["occ%d"%l] = p.occ[l].clone()
    for l in p.feats: d["feat%d"%l] = p.feats[l].clone()
    for l in p.warped: d["warped%d"%l] = p.warped[l].clone()
    for i,t in enumerate(p.output): d["out%02d"%i] = t.clone()
    return d
def cmp(a,b,tag):
    bad = [(k, float((a[k]-b[k]).abs().max())) for k in a if not bool((a[k]==b[k]).all())]
    print(tag, "differences:", bad)
net.forward(x, graph=False); e1 = snap()
net.forward(x, graph=False); e2 = snap()
cmp(e1,e2,"eager1 vs eager2")
net.forward(x, graph=True); g1 = snap()
net.forward(x, graph=True); g2 = snap()
cmp(e1,g1,"eager vs graph1")
cmp(g1,g2,"graph1 vs graph2")
# single-lane graph: all ops on lane 0
ops = p.ops
p.ops = [(0,)+op[1:] if op[0] != "fork" else op for op in ops]
p.graph = None
net.forward(x, graph=True); s1 = snap()
cmp(e1,s1,"eager vs single-lane graph")
net.forward(x, graph=False); s2 = snap()
cmp(e1,s2,"eager vs single-lane eager")
