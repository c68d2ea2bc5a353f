"""ctypes binding of oracle/_ref/libstn_ref.so -- the REFERENCE's own CUDA sampler.

TEST INFRASTRUCTURE ONLY (imported by tests/ and tests/golden/make_ref_sampler_golden.py; never by the
product).  The library is the reference's unmodified extras/stnbhwd/{utils.c,BilinearSamplerBHWD.cu}
(Lua-C glue :118-158, :313-419 and kernels :41-115, :161-307) compiled from where they lie under
/root/reference against the stand-in Torch7 headers of oracle/ref_shim/ (recipe: Makefile target
`oracle/_ref/libstn_ref.so`).  It needs a CUDA device to run, so it pins rows a4/a5 of SURVEY.md
section 8 on the GPU box, and -- through tests/golden/ref_sampler_golden.npz, which is its output -- the
numpy / C oracles on the CPU.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libstn_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.stn_ref_update_output.restype = C.c_int
        lib.stn_ref_update_output.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p]
        lib.stn_ref_update_grad_input.restype = C.c_int
        lib.stn_ref_update_grad_input.argtypes = [C.c_void_p] * 5 + [C.c_int] * 7 + [C.c_void_p]
        _lib = lib
    return _lib


def run(img, grid, grad_out, only_grid=False, device="cuda:0"):
    """BilinearSamplerBHWD.lua:53-115 around the reference's Lua-C entries: resize + zero-fill the
    results (:70, :99-102), call updateOutput / updateGradInput[OnlyGrid].  numpy in, numpy out:
    (output, gradImg or None, gradGrid)."""
    import torch
    lib = load()
    dev = torch.device(device)

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)

    ti, tg, tgo = t(img), t(grid), t(grad_out)
    B, H, W, Cn = ti.shape
    _, Hg, Wg, _ = tg.shape
    out = torch.zeros((B, Hg, Wg, Cn), device=dev)
    gi = torch.zeros_like(ti)
    gg = torch.zeros_like(tg)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda x: C.c_void_p(x.data_ptr())   # noqa: E731
    assert lib.stn_ref_update_output(p(ti), p(tg), p(out), B, H, W, Cn, Hg, Wg, st) == 0
    assert lib.stn_ref_update_grad_input(p(ti), p(tg), p(gi), p(gg), p(tgo), B, H, W, Cn, Hg, Wg,
                                         int(only_grid), st) == 0
    torch.cuda.synchronize(dev)
    return out.cpu().numpy(), (None if only_grid else gi.cpu().numpy()), gg.cpu().numpy()
