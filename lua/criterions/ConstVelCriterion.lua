-- Drop-in replacements of criterions/ConstVelCriterion.lua and criterions/OcclusionPriorCriterion.lua.
local b2f = require 'b2f_ffi'
local ffi = b2f.ffi

local ConstVel, parent = torch.class('nn.ConstVelCriterion', 'nn.Criterion')
function ConstVel:__init()
  parent.__init(self)
  self.sizeAverage = true
  self.gradCheck = false
end
function ConstVel:_run(input)
  assert(input[1]:nElement() == input[2]:nElement(), "input and target size mismatch")
  local f, b = input[1]:contiguous(), input[2]:contiguous()
  local gf, gb = f.new():resizeAs(f), b.new():resizeAs(b)
  local loss = ffi.new('double[1]')
  b2f.check(b2f.lib.b2f_constvel_criterion(f:data(), b:data(), f:size(1), f:size(2), f:size(3), f:size(4),
            self.sizeAverage and 1 or 0, gf:data(), gb:data(), nil, loss, b2f.stream()))
  self._grads = {gf, gb}
  return loss[0]
end
function ConstVel:updateOutput(input) self.output = self:_run(input); return self.output end
function ConstVel:updateGradInput(input)
  if not self._grads then self:_run(input) end
  local g = self._grads; self._grads = nil
  return g
end
function ConstVel:clear() self.output = nil; self.gradInput = nil; self._grads = nil end

local OccPrior, parent2 = torch.class('nn.OcclusionPriorCriterion', 'nn.Criterion')
function OccPrior:__init()
  parent2.__init(self)
  self.sizeAverage = true
  self.penalty = 1
end
function OccPrior:_run(input, target)
  assert(input:size(3) == target:size(3) and input:size(4) == target:size(4), "input and target size mismatch")
  local occ = input:contiguous()
  local grad = occ.new():resizeAs(occ)
  local loss = ffi.new('double[1]')
  b2f.check(b2f.lib.b2f_occprior_criterion(occ:data(), occ:size(1), occ:size(2), occ:size(3), occ:size(4),
            self.penalty, self.sizeAverage and 1 or 0, grad:data(), nil, loss, b2f.stream()))
  self._grad = grad
  return loss[0]
end
function OccPrior:updateOutput(input, target) return self:_run(input, target) end
function OccPrior:updateGradInput(input, target)
  if not self._grad then self:_run(input, target) end
  local g = self._grad; self._grad = nil
  return g
end
