-- Drop-in replacements of criterions/OBCCriterion.lua and criterions/OBGCCriterion.lua: one fused
-- kernel call computes the loss and every gradient.  Fields, defaults and the :clear() method are the
-- reference's; gradients are cached between updateOutput and updateGradInput of the same inputs.
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

  function Crit:_run(input, target)
    assert(#input >= 4, "expecting at least four inputs")
    assert(self.F == 3, "libb2f_cuda: only F = 3 is implemented")
    local ws = self.past_flow and 4 or 3
    local flow, occ, wp, wf = input[1], input[ws-1], input[ws], input[ws+1]
    local bflow = self.past_flow and input[2] or nil
    assert(wp:nElement() == target:nElement(), "input and target size mismatch")
    local kind, eps = b2f.penalty(self.p)
    local prm = ffi.new('b2f_ob_params', {gradient_terms, kind, eps, self.penalty_out, self.alpha,
                        self.beta, self.gamma, self.pwc_flow_scaling, self.past_flow and 1 or 0,
                        self.gradCheck and 1 or 0, self.sizeAverage and 1 or 0})
    local g = {occ.new():resizeAs(occ), wp.new():resizeAs(wp), wf.new():resizeAs(wf)}
    local loss = ffi.new('double[1]')
    local tgt = target:contiguous()
    b2f.check(b2f.lib.b2f_ob_criterion(prm, flow:data(), b2f.ptr(bflow), occ:data(), wp:data(), wf:data(),
              tgt:data(), tgt:size(1), tgt:size(2), tgt:size(3), tgt:size(4),
              g[1]:data(), g[2]:data(), g[3]:data(), nil, loss, b2f.stream()))
    self._grads = g
    return loss[0]
  end

  function Crit:updateOutput(input, target)
    self.output = self:_run(input, target)
    return self.output
  end

  function Crit:updateGradInput(input, target)
    if not self._grads then self:_run(input, target) end
    local g = self._grads
    self._grads = nil
    return g          -- fresh table {gradOcc, gradWarp_1, gradWarp_2} (OBCCriterion.lua:132-135)
  end

  function Crit:clear()
    self.coord = nil
    self._grads = nil
  end
end

define('nn.OBCCriterion', 0)
define('nn.OBGCCriterion', 1)
