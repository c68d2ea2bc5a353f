-- nn.BilinearSamplerBHWD on top of libb2f_cuda.so (LuaJIT FFI shim, see INTEGRATION.md).
--
-- Replaces the three Lua-C entries the reference registers on the CudaTensor metatable
-- (extras/stnbhwd/BilinearSamplerBHWD.cu:423-435) by b2f_warp_bhwd_forward / _backward.  Module
-- behaviour mirrors extras/stnbhwd/BilinearSamplerBHWD.lua: {images BHWD, grids BHW2} in, BHWD
-- out; 3-D inputs are treated as a batch of one; both gradInputs are freshly sized and the image
-- gradient zero-filled before the scatter; shape violations raise the same assertion errors.
local b2f = require 'b2f_ffi'
local Sampler, Base = torch.class('nn.BilinearSamplerBHWD', 'nn.Module')

function Sampler:__init()
  Base.__init(self)
  self.gradInput = {}
end

-- view a 3-D tensor as a batch of one (no copy)
local function batched(t)
  if t:nDimension() ~= 3 then return t end
  return t:view(1, t:size(1), t:size(2), t:size(3))
end

function Sampler:check(input, gradOutput)
  local images, grids = input[1], input[2]
  assert(images:isContiguous(), 'Input images have to be contiguous')
  assert(images:nDimension() == 4 and grids:nDimension() == 4)
  assert(images:size(1) == grids:size(1), 'batch size mismatch')
  assert(grids:size(4) == 2, 'grids need two coordinates (x offset, y offset)')
  if gradOutput then
    for d = 1, 3 do assert(grids:size(d) == gradOutput:size(d), 'gradOutput / grids size mismatch') end
  end
end

local function dims(images, grids)
  return images:size(1), images:size(2), images:size(3), images:size(4), grids:size(2), grids:size(3)
end

function Sampler:updateOutput(input)
  local single = input[1]:nDimension() == 3
  local images, grids = batched(input[1]), batched(input[2])
  self:check({images, grids})
  grids = grids:contiguous()
  local B, H, W, C, Hg, Wg = dims(images, grids)
  self.output:resize(B, Hg, Wg, C)
  b2f.check(b2f.lib.b2f_warp_bhwd_forward(images:data(), grids:data(), self.output:data(),
                                          B, H, W, C, Hg, Wg, b2f.stream()))
  if single then self.output = self.output:select(1, 1) end
  return self.output
end

function Sampler:updateGradInput(input, gradOutput)
  local single = input[1]:nDimension() == 3
  local images, grids, go = batched(input[1]), batched(input[2]), batched(gradOutput)
  self:check({images, grids}, go)
  grids, go = grids:contiguous(), go:contiguous()
  local B, H, W, C, Hg, Wg = dims(images, grids)
  local gImages = (self.gradInput[1] or images.new()):resizeAs(images)
  local gGrids = (self.gradInput[2] or images.new()):resizeAs(grids)
  local stream = b2f.stream()
  -- the native entry accumulates into the image gradient, exactly like the reference's kernel
  b2f.check(b2f.lib.b2f_zero_async(gImages:data(), gImages:nElement() * 4, stream))
  b2f.check(b2f.lib.b2f_warp_bhwd_backward(images:data(), grids:data(), go:data(), gImages:data(),
                                           gGrids:data(), B, H, W, C, Hg, Wg, stream))
  if single then gImages, gGrids = gImages:select(1, 1), gGrids:select(1, 1) end
  self.gradInput = {gImages, gGrids}
  return self.gradInput
end
