-- Drop-in replacements of criterions/ConstVelCriterion.lua and criterions/OcclusionPriorCriterion.lua.
local b2f = require 'b2f_ffi'
local ffi = b2f.ffi

local ConstVel, parent = torch.class('nn.ConstVelCriterion', 'nn.Criterion')
function ConstVel:__init()
  parent.__init(self)
  self.sizeAverage = true
  self.gradCheck = false
end
function ConstVel:_run(input, want_grads)
  assert(input[1]:nElement() == input[2]:nElement(), "input and target size mismatch")
  local f, b = input[1]:contiguous(), input[2]:contiguous()
  local gf, gb
  if want_grads then gf, gb = f.new():resizeAs(f), b.new():resizeAs(b) end
  local loss = ffi.new('double[1]')
  b2f.check(b2f.lib.b2f_constvel_criterion(f:data(), b:data(), f:size(1), f:size(2), f:size(3), f:size(4),
            self.sizeAverage and 1 or 0, b2f.ptr(gf), b2f.ptr(gb), nil, loss, b2f.stream()))
  return loss[0], {gf, gb}
end
-- updateGradInput recomputes like the reference (:48-74) unless self.fuse_backward (see OBCCriterion.lua here)
function ConstVel:updateOutput(input)
  local loss, g = self:_run(input, self.fuse_backward)
  self.output = loss
  self._held = self.fuse_backward and {input[1], input[2], self.sizeAverage, g} or nil
  return self.output
end
function ConstVel:updateGradInput(input)
  local held = self._held
  self._held = nil
  if held and self.fuse_backward and rawequal(held[1], input[1]) and rawequal(held[2], input[2])
     and held[3] == self.sizeAverage then return held[4] end
  local _, g = self:_run(input, true)
  return g
end
function ConstVel:clear() self.output = nil; self.gradInput = nil; self._held = nil end

local OccPrior, parent2 = torch.class('nn.OcclusionPriorCriterion', 'nn.Criterion')
function OccPrior:__init()
  parent2.__init(self)
  self.sizeAverage = true
  self.penalty = 1
end
function OccPrior:_run(input, target, want_grad)
  assert(input:size(3) == target:size(3) and input:size(4) == target:size(4), "input and target size mismatch")
  local occ = input:contiguous()
  local grad = want_grad and occ.new():resizeAs(occ) or nil
  local loss = ffi.new('double[1]')
  b2f.check(b2f.lib.b2f_occprior_criterion(occ:data(), occ:size(1), occ:size(2), occ:size(3), occ:size(4),
            self.penalty, self.sizeAverage and 1 or 0, b2f.ptr(grad), nil, loss, b2f.stream()))
  return loss[0], grad
end
function OccPrior:updateOutput(input, target)
  local loss, grad = self:_run(input, target, self.fuse_backward)
  self.output = loss
  self._held = self.fuse_backward and {input, self.penalty, self.sizeAverage, grad} or nil
  return self.output
end
function OccPrior:updateGradInput(input, target)
  local held = self._held
  self._held = nil
  if held and self.fuse_backward and rawequal(held[1], input) and held[2] == self.penalty
     and held[3] == self.sizeAverage then return held[4] end
  local _, grad = self:_run(input, target, true)
  return grad
end
