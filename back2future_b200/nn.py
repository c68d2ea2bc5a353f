"""Host-side mirror of the reference's Torch7 ``nn.Module`` / ``nn.Criterion`` surface for the hot path.

The reference's host language (Lua/Torch7) is not available in this environment, so the host side
above the C ABI is written in Python with the same class names, constructor signatures, fields,
method names, argument meaning and error behaviour as the reference's modules; ``lua/`` holds the
LuaJIT-FFI shims a maintainer would drop into the reference.  torch is used
ONLY as the owner of device memory and streams (the role cutorch plays in the reference); every
computation is a call into libb2f_cuda.so.  There is no CPU path: CPU tensors are rejected.

Reference files mirrored (relative to the reference root):
  models/CostVolMulti.lua, extras/stnbhwd/BilinearSamplerBHWD.lua, criterions/OBCCriterion.lua,
  criterions/OBGCCriterion.lua, criterions/SmoothnessCriterion.lua,
  criterions/SecondOrderSmoothnessCriterion.lua, criterions/ConstVelCriterion.lua,
  criterions/OcclusionPriorCriterion.lua, criterions/penalty/*.lua
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

__all__ = [
    "CostVolMulti", "BilinearSamplerBHWD", "OBCCriterion", "OBGCCriterion", "SmoothnessCriterion",
    "SecondOrderSmoothnessCriterion", "ConstVelCriterion", "OcclusionPriorCriterion",
    "QuadraticPenalty", "L1Penalty", "LorentzianPenalty",
]


# ---------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------

def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, what):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s: expected a tensor, got %r" % (what, type(t)))
    if not t.is_cuda:
        raise RuntimeError("%s: CPU tensor given; the B200 path has no CPU fallback" % what)
    if t.dtype != torch.float32:
        raise TypeError("%s: expected float32, got %s" % (what, t.dtype))
    return t


def _contig(t, what):
    """The C ABI takes dense tensors.  A strided view is REJECTED, not copied: `.contiguous()` would launch a torch
    kernel on the product path (north_star: torch owns memory and streams, it does not launch work).  The Lua shims
    call :contiguous() at the same places, i.e. there the copy is the caller's Torch7, as in the reference."""
    _dev(t, what)
    if not t.is_contiguous():
        raise ValueError("%s: non-contiguous tensor (strides %r for shape %r); pass a dense tensor -- this path does "
                         "not copy behind the caller's back" % (what, tuple(t.stride()), tuple(t.shape)))
    return t


def _zeros_like(t):
    """A zero-filled buffer without a torch kernel: torch allocates, b2f_zero_async (cudaMemsetAsync on the
    current stream) fills -- what `gradInput:zero()` is in the reference's wrapper (BilinearSamplerBHWD.lua:99-102)."""
    z = torch.empty_like(t)
    _lib.check(_lib.load().b2f_zero_async(_p(z), z.numel() * 4, _stream()))
    return z


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


# ---------------------------------------------------------------------------------------
# penalties (criterions/penalty/*.lua) -- carriers of the enum the kernels switch on
# ---------------------------------------------------------------------------------------

class QuadraticPenalty:
    """quadratic_function.lua:15-21."""
    kind = _lib.PENALTY_QUADRATIC
    eps = 0.0


class L1Penalty:
    """L1_function.lua:15-26.  ``alpha`` is accepted and ignored exactly like the reference
    (``self.alpha = 0.5 or alpha``)."""
    kind = _lib.PENALTY_L1

    def __init__(self, alpha=None):
        self.eps = 0.001 * 0.001
        self.alpha = 0.5


class LorentzianPenalty:
    """Lorentzian_function.lua:15-31."""
    kind = _lib.PENALTY_LORENTZIAN

    def __init__(self):
        self.set_eps(0.05)

    def set_eps(self, eps):
        self.eps = eps
        self.eps_sq = eps * eps


def _pen(p):
    kind = getattr(p, "kind", None)
    if kind not in (0, 1, 2):
        raise TypeError("unsupported penalty object %r" % (p,))
    eps = float(p.eps) if kind == _lib.PENALTY_LORENTZIAN else 0.0
    return kind, eps


# ---------------------------------------------------------------------------------------
# nn.Module protocol
# ---------------------------------------------------------------------------------------

class Module:
    """The slice of Torch7's nn.Module the hot-path modules use."""

    def __init__(self):
        self.output = torch.empty(0)
        self.gradInput = torch.empty(0)
        self.train = True

    def updateOutput(self, input):
        raise NotImplementedError

    def updateGradInput(self, input, gradOutput):
        raise NotImplementedError

    def accGradParameters(self, input, gradOutput, scale=1):
        """No parameters anywhere on the hot path."""
        return None

    def forward(self, input):
        return self.updateOutput(input)

    def backward(self, input, gradOutput, scale=1):
        g = self.updateGradInput(input, gradOutput)
        self.accGradParameters(input, gradOutput, scale)
        return g

    def parameters(self):
        return None

    def zeroGradParameters(self):
        return None

    def training(self):
        self.train = True
        return self

    def evaluate(self):
        self.train = False
        return self

    def cuda(self):
        return self

    def type(self, _t=None):
        return self

    def clearState(self):
        self.output = torch.empty(0)
        if isinstance(self.gradInput, (list, tuple)):
            self.gradInput = [torch.empty(0) for _ in self.gradInput]
        else:
            self.gradInput = torch.empty(0)
        return self


class CostVolMulti(Module):
    """nn.CostVolMulti(win, fwd, verbose) -- models/CostVolMulti.lua:23-47.

    ``input`` is a list {ref, frame_2, ..., frame_F} of (B,C,h,w) maps; the output has win*win
    channels, x-major."""

    def __init__(self, win=None, fwd=None, verbose=None):
        super().__init__()
        self.win = win if win else 3
        self.fwd = fwd if fwd is not None else True
        self.verbose = verbose if verbose else False
        self.gradInput = [torch.empty(0), torch.empty(0)]

    def _check(self, input):
        frames = [_contig(t, "CostVolMulti input[%d]" % (i + 1)) for i, t in enumerate(input)]
        for f in range(1, len(frames)):
            assert frames[f].numel() == frames[f - 1].numel(), "input sizes mismatch"
        if frames[0].dim() != 4:
            raise RuntimeError("CostVolMulti: expected 4-D (B,C,h,w) inputs")
        return frames

    def updateOutput(self, input, out=None):
        """CostVolMulti.lua:49-109.  ``out`` (optional, beyond the reference) lets the caller
        supply a (B, win*win, h, w) view of a wider buffer -- e.g. one half of the 162-channel
        JoinTable output -- whose batch stride is then passed through the ABI."""
        lib = _lib.load()
        frames = self._check(input)
        B, Cn, h, w = frames[0].shape
        ww = self.win * self.win
        if out is None:
            if self.output.shape != (B, ww, h, w) or self.output.device != frames[0].device:
                self.output = torch.empty((B, ww, h, w), device=frames[0].device, dtype=torch.float32)
            out = self.output
        else:
            _dev(out, "CostVolMulti out")
            assert tuple(out.shape) == (B, ww, h, w) and out.stride()[1:] == (h * w, w, 1)
            self.output = out
        obs = out.stride(0) if B > 1 else ww * h * w
        ptrs = _lib.ptr_array([t.data_ptr() for t in frames])
        _lib.check(lib.b2f_costvol_forward(ptrs, len(frames), B, Cn, h, w, int(self.win), int(bool(self.fwd)),
                                           _p(out), obs, _stream()))
        if self.verbose:
            print(tuple(self.output.shape))
        return self.output

    def updateGradInput(self, input, gradOutput):
        """CostVolMulti.lua:111-181.  ``gradOutput`` may be a narrow of a wider buffer (batch-
        strided); anything else non-contiguous is rejected."""
        lib = _lib.load()
        frames = self._check(input)
        B, Cn, h, w = frames[0].shape
        ww = self.win * self.win
        go = _dev(gradOutput, "CostVolMulti gradOutput")
        assert tuple(go.shape) == (B, ww, h, w), "gradOutput size mismatch"
        if go.stride()[1:] != (h * w, w, 1):
            raise ValueError("CostVolMulti gradOutput: only a batch-strided (B, win*win, h, w) view is accepted, got "
                             "strides %r" % (tuple(go.stride()),))
        gbs = go.stride(0) if B > 1 else ww * h * w
        if len(self.gradInput) != len(frames):
            self.gradInput = [torch.empty(0) for _ in frames]
        for f, t in enumerate(frames):
            if self.gradInput[f].shape != t.shape or self.gradInput[f].device != t.device:
                self.gradInput[f] = torch.empty_like(t)
        fptrs = _lib.ptr_array([t.data_ptr() for t in frames])
        gptrs = _lib.ptr_array([t.data_ptr() for t in self.gradInput])
        _lib.check(lib.b2f_costvol_backward(fptrs, len(frames), B, Cn, h, w, int(self.win), int(bool(self.fwd)),
                                            _p(go), gbs, gptrs, _stream()))
        return self.gradInput

    def __repr__(self):
        return "nn.CostVolMulti" + "window size = %d" % self.win


class BilinearSamplerBHWD(Module):
    """nn.BilinearSamplerBHWD() -- extras/stnbhwd/BilinearSamplerBHWD.lua.

    input = {inputImages (B,H,W,C) contiguous, grids (B,Hg,Wg,2)}; 3-D inputs get an outer
    batch dimension (:58-64).  Semantics are those of the reference's CUDA kernel."""

    def __init__(self):
        super().__init__()
        self.gradInput = []

    @staticmethod
    def check(input, gradOutput=None):
        """BilinearSamplerBHWD.lua:26-41."""
        inputImages, grids = input
        assert inputImages.is_contiguous(), "Input images have to be contiguous"
        assert inputImages.dim() == 4
        assert grids.dim() == 4
        assert inputImages.size(0) == grids.size(0)
        assert grids.size(3) == 2
        if gradOutput is not None:
            assert grids.size(0) == gradOutput.size(0)
            assert grids.size(1) == gradOutput.size(1)
            assert grids.size(2) == gradOutput.size(2)

    def updateOutput(self, input):
        lib = _lib.load()
        _img, _grid = _dev(input[0], "inputImages"), _dev(input[1], "grids")
        squeeze = _img.dim() == 3
        img = _img.unsqueeze(0) if squeeze else _img
        grid = _grid.unsqueeze(0) if squeeze else _grid
        self.check((img, grid))
        grid = _contig(grid, "grids")
        B, H, W, Cn = img.shape
        _, Hg, Wg, _ = grid.shape
        if self.output.shape != (B, Hg, Wg, Cn) or self.output.device != img.device:
            self.output = torch.empty((B, Hg, Wg, Cn), device=img.device, dtype=torch.float32)
        _lib.check(lib.b2f_warp_bhwd_forward(_p(img), _p(grid), _p(self.output), B, H, W, Cn, Hg, Wg, _stream()))
        if squeeze:
            self.output = self.output[0]
        return self.output

    def updateGradInput(self, input, gradOutput, only_grid=False):
        """BilinearSamplerBHWD.lua:81-115: both gradInputs are zero-filled, then one kernel.
        ``only_grid=True`` maps to the reference's registered-but-unused
        BilinearSamplerBHWD_updateGradInputOnlyGrid (.cu:368-419): gradInput[1] is left zero."""
        lib = _lib.load()
        _img, _grid, _go = _dev(input[0], "inputImages"), _dev(input[1], "grids"), _dev(gradOutput, "gradOutput")
        squeeze = _img.dim() == 3
        img = _img.unsqueeze(0) if squeeze else _img
        grid = _grid.unsqueeze(0) if squeeze else _grid
        go = _go.unsqueeze(0) if squeeze else _go
        self.check((img, grid), go)
        grid = _contig(grid, "grids")
        go = _contig(go, "gradOutput")
        B, H, W, Cn = img.shape
        _, Hg, Wg, _ = grid.shape
        gimg = _zeros_like(img)
        ggrid = torch.empty_like(grid)
        _lib.check(lib.b2f_warp_bhwd_backward(_p(img), _p(grid), _p(go), None if only_grid else _p(gimg),
                                              _p(ggrid), B, H, W, Cn, Hg, Wg, _stream()))
        self.gradInput = [gimg[0], ggrid[0]] if squeeze else [gimg, ggrid]
        return self.gradInput


class WarpingUnit(Module):
    """`warpingUnit(I, F)` of models/pwc.lua:68-73 as ONE module, with the `nn.MulConstant(flow_scale)` that feeds
    it (:402-408 feature warps, :441-446 image warps) folded in: input = {I (B,C,H,W), F (B,2,H,W)}, both in the
    network's own BDHW layout; output (B,C,H,W).  Replaces Transpose x2 -> BilinearSamplerBHWD -> Transpose and the
    scaling pass; results equal that chain's."""

    def __init__(self, flow_scale=1.0):
        super().__init__()
        self.flow_scale = float(flow_scale)
        self.gradInput = []

    @staticmethod
    def check(input, gradOutput=None):
        I, F = input
        assert I.dim() == 4 and F.dim() == 4
        assert I.size(0) == F.size(0) and F.size(1) == 2
        assert I.size(2) == F.size(2) and I.size(3) == F.size(3)
        if gradOutput is not None:
            assert tuple(gradOutput.shape) == tuple(I.shape)

    def updateOutput(self, input):
        lib = _lib.load()
        I, F = _dev(input[0], "I"), _dev(input[1], "F")
        self.check((I, F))
        I, F = _contig(I, "I"), _contig(F, "F")
        B, Cn, H, W = I.shape
        if self.output.shape != I.shape or self.output.device != I.device:
            self.output = torch.empty_like(I)
        _lib.check(lib.b2f_warp_bdhw_forward(_p(I), _p(F), self.flow_scale, _p(self.output), B, Cn, H, W, _stream()))
        return self.output

    def updateGradInput(self, input, gradOutput, only_grid=False):
        lib = _lib.load()
        I, F, go = _dev(input[0], "I"), _dev(input[1], "F"), _dev(gradOutput, "gradOutput")
        self.check((I, F), go)
        I, F, go = _contig(I, "I"), _contig(F, "F"), _contig(go, "gradOutput")
        B, Cn, H, W = I.shape
        gI = _zeros_like(I)
        gF = torch.empty_like(F)
        _lib.check(lib.b2f_warp_bdhw_backward(_p(I), _p(F), self.flow_scale, _p(go), None if only_grid else _p(gI),
                                              _p(gF), B, Cn, H, W, _stream()))
        self.gradInput = [gI, gF]
        return self.gradInput


# ---------------------------------------------------------------------------------------
# nn.Criterion protocol
# ---------------------------------------------------------------------------------------

class Criterion:
    """nn.Criterion protocol.

    The reference's criterions recompute everything in updateGradInput (e.g. OBCCriterion.lua:121-240), and so do
    these by default: updateOutput runs the fused kernel without gradient outputs (loss only), updateGradInput runs
    it again with them.  Nothing computed in the forward is handed out later, so a buffer that was rewritten in
    between -- through this library's raw pointers, invisibly to torch's version counters -- is always seen.

    `fuse_backward = True` (an extension, off by default) makes updateOutput produce the gradients in the same
    pass and updateGradInput hand them out, IF it is called with the very same objects (`is`) and unchanged
    criterion fields; the caller then guarantees that the buffers are not rewritten between the two calls, which is
    train.lua's pattern (forward at :428-444 immediately followed by backward at :445-475)."""

    fuse_backward = False

    def __init__(self):
        self.output = 0
        self.gradInput = torch.empty(0)
        self._held = None       # (objects, params, grads) of the last fused forward

    def forward(self, input, target=None):
        return self.updateOutput(input, target)

    def backward(self, input, target=None):
        return self.updateGradInput(input, target)

    def cuda(self):
        return self

    def type(self, _t=None):
        return self

    def clear(self):
        """The non-standard :clear() train.lua calls after each level (train.lua:433, 454, 461)."""
        self._held = None

    def _params(self):
        return ()

    def _hold(self, objects, grads):
        self._held = (tuple(objects), self._params(), grads) if self.fuse_backward else None

    def _take(self, objects):
        """Gradients of the fused forward if it saw exactly these objects and fields; None otherwise."""
        held, self._held = self._held, None
        if held is None or not self.fuse_backward:
            return None
        objs, params, grads = held
        objects = tuple(objects)
        if len(objs) != len(objects) or any(a is not b for a, b in zip(objs, objects)) or params != self._params():
            return None
        return grads


def _loss_call(fn, *args):
    """Call a criterion entry with a host double for the loss (synchronises, like the reference's
    criterions which return a Lua number)."""
    host = C.c_double(0.0)
    _lib.check(fn(*args, None, C.byref(host), _stream()))
    return host.value


class _OBBase(Criterion):
    _gradient_terms = 0

    def __init__(self):
        super().__init__()
        self.sizeAverage = True
        self.gradCheck = False
        self.p = QuadraticPenalty()
        self.penalty_out = 1.0
        self.F = 3
        self.pwc_flow_scaling = 1
        self.past_flow = False
        self.alpha = 1.0
        self.beta = 1.0
        self.gamma = 1.0

    def _params(self):
        return (self._gradient_terms, _pen(self.p), float(self.penalty_out), float(self.alpha), float(self.beta),
                float(self.gamma), float(self.pwc_flow_scaling), bool(self.past_flow), bool(self.gradCheck),
                bool(self.sizeAverage), int(self.F))

    def _objects(self, input, target):
        warp_start = 3 if self.past_flow else 2  # 0-based index of the first warped frame
        return [input[0], input[1] if self.past_flow else None, input[warp_start - 1], input[warp_start],
                input[warp_start + 1], target]

    def _run(self, input, target, want_grads):
        lib = _lib.load()
        assert len(input) >= 4, "expecting at least four inputs"
        if self.F != 3:
            raise _lib.B2FError(-2, "OBCC/OBGCC: only F = 3 (two warped frames) is implemented, got F=%r" % self.F)
        o_flow, o_bflow, o_occ, o_wp, o_wf, o_tgt = self._objects(input, target)
        flow = _contig(o_flow, "flow")
        bflow = _contig(o_bflow, "bflow") if self.past_flow else None
        occ = _contig(o_occ, "occ")
        wp = _contig(o_wp, "warped frame 1")
        wf = _contig(o_wf, "warped frame 2")
        tgt = _contig(o_tgt, "target")
        assert wp.numel() == tgt.numel() and wf.numel() == tgt.numel(), "input and target size mismatch"
        B, Cn, h, w = tgt.shape
        kind, eps = _pen(self.p)
        prm = _lib.ObParams(self._gradient_terms, kind, eps, float(self.penalty_out), float(self.alpha),
                            float(self.beta), float(self.gamma), float(self.pwc_flow_scaling),
                            int(bool(self.past_flow)), int(bool(self.gradCheck)), int(bool(self.sizeAverage)))
        grads = [torch.empty_like(occ), torch.empty_like(wp), torch.empty_like(wf)] if want_grads else [None] * 3
        loss = _loss_call(lib.b2f_ob_criterion, C.byref(prm), _p(flow), _p(bflow), _p(occ), _p(wp), _p(wf),
                          _p(tgt), B, Cn, h, w, _p(grads[0]), _p(grads[1]), _p(grads[2]))
        return loss, grads

    def updateOutput(self, input, target):
        self.output, grads = self._run(input, target, self.fuse_backward)
        self._hold(self._objects(input, target), grads)
        return self.output

    def updateGradInput(self, input, target):
        grads = self._take(self._objects(input, target))
        if grads is None:
            _, grads = self._run(input, target, True)
        return grads  # fresh table {gradOcc, gradWarp_1, gradWarp_2} (OBCCriterion.lua:132-135)


class OBCCriterion(_OBBase):
    """nn.OBCCriterion -- criterions/OBCCriterion.lua.  input = {flow, [bflow], occ, warp_1, warp_2}."""
    _gradient_terms = 0


class OBGCCriterion(_OBBase):
    """nn.OBGCCriterion -- criterions/OBGCCriterion.lua (alpha/beta/gamma; alpha is backward-only, Q5)."""
    _gradient_terms = 1


class _SmoothBase(Criterion):
    _order = 1

    def __init__(self):
        super().__init__()
        self.sizeAverage = True
        self.gradCheck = False
        self.p = QuadraticPenalty()
        self.cs = 20
        # parity default: reproduce the Torch7 view-resize aliasing of the edge weights
        self.alias_weights = True

    def _params(self):
        return (self._order, _pen(self.p), float(self.cs), bool(self.sizeAverage), bool(self.alias_weights))

    def _run(self, input, target, want_grads):
        lib = _lib.load()
        inp = _contig(input, "input")
        tgt = _contig(target, "target")
        assert inp.size(2) == tgt.size(2) and inp.size(3) == tgt.size(3), "input and target size mismatch"
        B, Cin, h, w = inp.shape
        kind, eps = _pen(self.p)
        prm = _lib.SmoothParams(self._order, kind, eps, float(self.cs), int(bool(self.sizeAverage)),
                                int(bool(self.alias_weights)))
        grad = torch.empty_like(inp) if want_grads else None
        loss = _loss_call(lib.b2f_smoothness_criterion, C.byref(prm), _p(inp), _p(tgt), B, Cin, tgt.size(1),
                          h, w, _p(grad))
        return loss, grad

    def updateOutput(self, input, target):
        self.output, grad = self._run(input, target, self.fuse_backward)
        self._hold([input, target], grad)
        return self.output

    def updateGradInput(self, input, target):
        grad = self._take([input, target])
        if grad is None:
            _, grad = self._run(input, target, True)
        return grad  # fresh tensor, not self.gradInput (Q10)


class SmoothnessCriterion(_SmoothBase):
    """nn.SmoothnessCriterion -- criterions/SmoothnessCriterion.lua."""
    _order = 1


class SecondOrderSmoothnessCriterion(_SmoothBase):
    """nn.SecondOrderSmoothnessCriterion -- criterions/SecondOrderSmoothnessCriterion.lua."""
    _order = 2


class ConstVelCriterion(Criterion):
    """nn.ConstVelCriterion -- criterions/ConstVelCriterion.lua.  Called with the whole output table;
    input[1], input[2] are the future and past flow (train.lua:437-440)."""

    def __init__(self):
        super().__init__()
        self.sizeAverage = True
        self.gradCheck = False

    def _params(self):
        return (bool(self.sizeAverage),)

    def _run(self, input, want_grads):
        lib = _lib.load()
        f = _contig(input[0], "input[1]")
        b = _contig(input[1], "input[2]")
        assert f.numel() == b.numel(), "input and target size mismatch"
        B, Cn, h, w = f.shape
        grads = [torch.empty_like(f), torch.empty_like(b)] if want_grads else [None, None]
        loss = _loss_call(lib.b2f_constvel_criterion, _p(f), _p(b), B, Cn, h, w, int(bool(self.sizeAverage)),
                          _p(grads[0]), _p(grads[1]))
        return loss, grads

    def updateOutput(self, input, target=None):
        self.output, grads = self._run(input, self.fuse_backward)
        self._hold([input[0], input[1]], grads)
        return self.output

    def updateGradInput(self, input, target=None):
        grads = self._take([input[0], input[1]])
        if grads is None:
            _, grads = self._run(input, True)
        return grads


class OcclusionPriorCriterion(Criterion):
    """nn.OcclusionPriorCriterion -- criterions/OcclusionPriorCriterion.lua (2- and 3-channel maps)."""

    def __init__(self):
        super().__init__()
        self.sizeAverage = True
        self.penalty = 1

    def _params(self):
        return (float(self.penalty), bool(self.sizeAverage))

    def _run(self, input, target, want_grads):
        lib = _lib.load()
        occ = _contig(input, "input")
        if target is not None:
            assert occ.size(2) == target.size(2) and occ.size(3) == target.size(3), "input and target size mismatch"
        B, Cn, h, w = occ.shape
        grad = torch.empty_like(occ) if want_grads else None
        loss = _loss_call(lib.b2f_occprior_criterion, _p(occ), B, Cn, h, w, float(self.penalty),
                          int(bool(self.sizeAverage)), _p(grad))
        return loss, grad

    def updateOutput(self, input, target=None):
        self.output, grad = self._run(input, target, self.fuse_backward)
        self._hold([input], grad)
        return self.output

    def updateGradInput(self, input, target=None):
        grad = self._take([input])
        if grad is None:
            _, grad = self._run(input, target, True)
        return grad
