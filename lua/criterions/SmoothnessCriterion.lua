-- Drop-in replacements of criterions/SmoothnessCriterion.lua and SecondOrderSmoothnessCriterion.lua.
require 'criterions.penalty.quadratic_function'
local b2f = require 'b2f_ffi'
local ffi = b2f.ffi

local function define(name, order)
  local Crit, parent = torch.class(name, 'nn.Criterion')

  function Crit:__init()
    parent.__init(self)
    self.sizeAverage = true
    self.gradCheck = false
    self.p = QuadraticPenalty()
    self.cs = 20
    self.alias_weights = true   -- reproduce the view-resize aliasing of :49-59 (SURVEY Q9)
  end

  function Crit:_run(input, target)
    assert(input:size(3) == target:size(3) and input:size(4) == target:size(4), "input and target size mismatch")
    local kind, eps = b2f.penalty(self.p)
    local prm = ffi.new('b2f_smooth_params', {order, kind, eps, self.cs, self.sizeAverage and 1 or 0,
                        self.alias_weights and 1 or 0})
    local inp, tgt = input:contiguous(), target:contiguous()
    local grad = inp.new():resizeAs(inp)
    local loss = ffi.new('double[1]')
    b2f.check(b2f.lib.b2f_smoothness_criterion(prm, inp:data(), tgt:data(), inp:size(1), inp:size(2), tgt:size(2),
              inp:size(3), inp:size(4), grad:data(), nil, loss, b2f.stream()))
    self._grad = grad
    return loss[0]
  end

  function Crit:updateOutput(input, target)
    self.output = self:_run(input, target)
    return self.output
  end

  function Crit:updateGradInput(input, target)
    if not self._grad then self:_run(input, target) end
    local g = self._grad
    self._grad = nil
    return g          -- a fresh tensor, like the reference (Q10)
  end

  function Crit:clear()
    self.buffer, self.gy, self.gx, self.wy, self.wx, self._grad = nil, nil, nil, nil, nil, nil
  end
end

define('nn.SmoothnessCriterion', 1)
define('nn.SecondOrderSmoothnessCriterion', 2)
