"""clock64 timeline of the tensor-core convolution's CTAs (b2f_debug_tc_trace): where a tile's residency goes.
usage: python tools/tc_trace.py [B=8]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from back2future_b200 import _lib

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Cin = Cout = 128
H, W = 112, 256
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
xh, xl = torch.randn(B, H, W, Cin, device="cuda"), torch.randn(B, H, W, Cin, device="cuda") * 1e-4
n = int(lib.b2f_conv3x3_tc_packed_floats(Cin, Cout))
wh, wl = torch.randn(n, device="cuda") * 0.03, torch.randn(n, device="cuda") * 1e-5
oh, ol = torch.empty(B, H, W, Cout, device="cuda"), torch.empty(B, H, W, Cout, device="cuda")
bias = torch.zeros(Cout, device="cuda")
run = lambda: _lib.check(lib.b2f_conv3x3_tc_forward(p(xh), p(xl), p(wh), p(wl), p(bias), p(oh), p(ol), None, 0, B, Cin, H, W, Cout, 0.2, st))
for _ in range(3):
    run()
ncta = 16 * 16 * B
buf = torch.zeros(ncta * 16, device="cuda", dtype=torch.int64)
lib.b2f_debug_tc_trace(p(buf))
run()
torch.cuda.synchronize()
lib.b2f_debug_tc_trace(None)
t = buf.cpu().numpy().reshape(ncta, 16).astype(np.int64)
names = ["entry->tmem ready", "tmem ready->first patch", "first patch->first weights", "first weights->last MMA issued",
         "last MMA issued->accumulator complete", "accumulator->tile staged", "staged->stores read", "stores->exit"]
d = np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4], t[:, 6] - t[:, 5],
              t[:, 7] - t[:, 6], t[:, 8] - t[:, 7]], 1)
tot = t[:, 8] - t[:, 0]
print("CTAs %d; residency median %d cycles (p10 %d, p90 %d); MMA floor of a tile = 4 chunks x 9 taps x 12 MMAs x 64 = 27648"
      % (ncta, np.median(tot), np.percentile(tot, 10), np.percentile(tot, 90)))
for i, nm in enumerate(names):
    print("  %-42s median %7d  p90 %7d" % (nm, np.median(d[:, i]), np.percentile(d[:, i], 90)))
# gap between consecutive CTAs on one SM
gaps = []
for sm in np.unique(t[:, 9]):
    rows = t[t[:, 9] == sm]
    rows = rows[np.argsort(rows[:, 0])]
    gaps += list(rows[1:, 0] - rows[:-1, 8])
print("gap between a CTA's exit and the next CTA's entry on the same SM: median %d, p90 %d cycles" % (np.median(gaps), np.percentile(gaps, 90)))
