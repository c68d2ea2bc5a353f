import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from back2future_b200 import _lib
from oracle import b2f_oracle as o
lib = _lib.load()
which = sys.argv[1]
B,Cn,h,w = 2,32,16,64
r = np.random.default_rng(4)
frames = [r.standard_normal((B,Cn,h,w)).astype(np.float32) for _ in range(2)]
wide = r.standard_normal((B,162,h,w)).astype(np.float32)
ft = [torch.from_numpy(f).cuda() for f in frames]; gw = torch.from_numpy(wide).cuda()
go = gw[:, :81]
grads = [torch.zeros_like(ft[0]), torch.zeros_like(ft[1])]
gp = [grads[0].data_ptr() if which in ('ref','both') else None, grads[1].data_ptr() if which in ('frm','both') else None]
lib.b2f_debug_costvol_path(2)
rc = lib.b2f_costvol_backward(_lib.ptr_array([t.data_ptr() for t in ft]), 2, B, Cn, h, w, 9, 1, C.c_void_p(go.data_ptr()), gw.stride(0), _lib.ptr_array(gp), None)
print('rc', rc, lib.b2f_last_error())
torch.cuda.synchronize()
ref = o.costvol_backward(frames, wide[:, :81], 9, True)
if which in ('ref','both'): print('gradRef err', o.rel_err(grads[0].cpu().numpy(), ref[0]))
if which in ('frm','both'): print('gradFrame err', o.rel_err(grads[1].cpu().numpy(), ref[1]))
