import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from back2future_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0"); B = 8; P = lambda t: C.c_void_p(t.data_ptr())
h, w = 320, 640
flow, bflow = torch.randn(B, 2, h, w, device=dev) * 0.2, torch.randn(B, 2, h, w, device=dev) * 0.2
occ = torch.softmax(torch.randn(B, 2, h, w, device=dev), 1).contiguous()
w1, w2, tgt = (torch.rand(B, 3, h, w, device=dev) * 4.7 - 2.1 for _ in range(3))
g2a, g2b, g3a, g3b = torch.empty_like(flow), torch.empty_like(flow), torch.empty_like(w1), torch.empty_like(w1)
loss = torch.zeros(1, dtype=torch.float64, device=dev)
for rep in range(2):
    for gt in (0, 1):
        prm = _lib.ObParams(gt, 1, 0.05, 1.0, 0.0 if gt else 1.0, 1.0, 1.0, 20.0, gt, 0, 0)
        _lib.check(lib.b2f_ob_criterion(C.byref(prm), P(flow), P(bflow), P(occ), P(w1), P(w2), P(tgt), B, 3, h, w, P(g2a), P(g3a), P(g3b), P(loss), None, None))
    for order in (1, 2):
        prm = _lib.SmoothParams(order, 1, 0.05, 20.0, 0, 1)
        _lib.check(lib.b2f_smoothness_criterion(C.byref(prm), P(flow), P(tgt), B, 2, 3, h, w, P(g2a), P(loss), None, None))
    _lib.check(lib.b2f_constvel_criterion(P(flow), P(bflow), B, 2, h, w, 1, P(g2a), P(g2b), P(loss), None, None))
    _lib.check(lib.b2f_occprior_criterion(P(occ), B, 2, h, w, 1.0, 0, P(g2a), P(loss), None, None))
torch.cuda.synchronize()
