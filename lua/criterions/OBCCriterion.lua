-- Drop-in replacements of criterions/OBCCriterion.lua and criterions/OBGCCriterion.lua: one fused
-- kernel call computes the loss and every gradient.  Fields, defaults and the :clear() method are the
-- reference's; updateGradInput recomputes like the reference unless self.fuse_backward is set.
require 'criterions.penalty.quadratic_function'
local b2f = require 'b2f_ffi'
local ffi = b2f.ffi

local function define(name, gradient_terms)
  local Crit, parent = torch.class(name, 'nn.Criterion')

  function Crit:__init()
    parent.__init(self)
    self.sizeAverage = true
    self.gradCheck = false
    self.p = QuadraticPenalty()
    self.penalty_out = 1.0
    self.alpha, self.beta, self.gamma = 1.0, 1.0, 1.0
    self.F = 3
    self.pwc_flow_scaling = 1
    self.past_flow = false
  end

  -- the objects a fused forward saw: updateGradInput hands its gradients out only for the very same ones
  local function objects(self, input, target)
    local ws = self.past_flow and 4 or 3
    return {input[1], self.past_flow and input[2] or false, input[ws-1], input[ws], input[ws+1], target}
  end
  local function fields(self)
    local kind, eps = b2f.penalty(self.p)
    return {kind, eps, self.penalty_out, self.alpha, self.beta, self.gamma, self.pwc_flow_scaling,
            self.past_flow, self.gradCheck, self.sizeAverage, self.F}
  end
  local function same(a, b)
    if not a or not b or #a ~= #b then return false end
    for i = 1, #a do if not rawequal(a[i], b[i]) and a[i] ~= b[i] then return false end end
    return true
  end

  function Crit:_run(input, target, want_grads)
    assert(#input >= 4, "expecting at least four inputs")
    assert(self.F == 3, "libb2f_cuda: only F = 3 is implemented")
    local ws = self.past_flow and 4 or 3
    -- the C ABI takes dense tensors: any strided model output (a narrow / select of a joined tensor) is made
    -- contiguous here by Torch7, exactly where the reference's own tensor ops would have handled the strides
    local flow, occ = input[1]:contiguous(), input[ws-1]:contiguous()
    local wp, wf = input[ws]:contiguous(), input[ws+1]:contiguous()
    local bflow = self.past_flow and input[2]:contiguous() or nil
    local tgt = target:contiguous()
    assert(wp:nElement() == tgt:nElement() and wf:nElement() == tgt:nElement(), "input and target size mismatch")
    local kind, eps = b2f.penalty(self.p)
    local prm = ffi.new('b2f_ob_params', {gradient_terms, kind, eps, self.penalty_out, self.alpha,
                        self.beta, self.gamma, self.pwc_flow_scaling, self.past_flow and 1 or 0,
                        self.gradCheck and 1 or 0, self.sizeAverage and 1 or 0})
    local g = want_grads and {occ.new():resizeAs(occ), wp.new():resizeAs(wp), wf.new():resizeAs(wf)} or {}
    local loss = ffi.new('double[1]')
    b2f.check(b2f.lib.b2f_ob_criterion(prm, flow:data(), b2f.ptr(bflow), occ:data(), wp:data(), wf:data(),
              tgt:data(), tgt:size(1), tgt:size(2), tgt:size(3), tgt:size(4),
              b2f.ptr(g[1]), b2f.ptr(g[2]), b2f.ptr(g[3]), nil, loss, b2f.stream()))
    return loss[0], g
  end

  -- Like the reference (OBCCriterion.lua:121-240) updateGradInput recomputes: the forward is a loss-only pass.
  -- self.fuse_backward = true (extension) lets the forward also produce the gradients, handed out by the next
  -- updateGradInput if it gets the very same tensors and unchanged fields (train.lua:428-475's pattern); the
  -- caller then guarantees the buffers are not rewritten in between.
  function Crit:updateOutput(input, target)
    local loss, g = self:_run(input, target, self.fuse_backward)
    self.output = loss
    self._held = self.fuse_backward and {objects(self, input, target), fields(self), g} or nil
    return self.output
  end

  function Crit:updateGradInput(input, target)
    local held = self._held
    self._held = nil
    if held and self.fuse_backward and same(held[1], objects(self, input, target)) and same(held[2], fields(self)) then
      return held[3]
    end
    local _, g = self:_run(input, target, true)
    return g          -- fresh table {gradOcc, gradWarp_1, gradWarp_2} (OBCCriterion.lua:132-135)
  end

  function Crit:clear()
    self.coord = nil
    self._held = nil
  end
end

define('nn.OBCCriterion', 0)
define('nn.OBGCCriterion', 1)
