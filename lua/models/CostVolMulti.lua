-- Drop-in replacement of models/CostVolMulti.lua: same class name, constructor and fields; the 81-pass
-- Lua loops (CostVolMulti.lua:59-100, 127-178) become one call each into libb2f_cuda.so.
require 'nn'
require 'cutorch'
local b2f = require 'b2f_ffi'
local ffi = b2f.ffi

local CostVolMulti, parent = torch.class('nn.CostVolMulti', 'nn.Module')

function CostVolMulti:__init(win, fwd, verbose)
  parent.__init(self)
  self.win = win or 3
  if fwd ~= nil then self.fwd = fwd else self.fwd = true end
  self.verbose = verbose or false
  self.gradInput = {torch.Tensor(), torch.Tensor()}
end

local function frame_ptrs(input)
  local frames = #input
  local arr = ffi.new('const float*[?]', frames)
  for f = 1, frames do
    assert(input[f]:isContiguous(), 'inputs have to be contiguous')
    arr[f-1] = input[f]:data()
  end
  return arr, frames
end

function CostVolMulti:updateOutput(input)
  for f = 2, #input do
    assert(input[f]:nElement() == input[f-1]:nElement(), "input sizes mismatch")
  end
  local ref = input[1]
  local B, N, h, w = ref:size(1), ref:size(2), ref:size(3), ref:size(4)
  self.output:resize(B, self.win * self.win, h, w)
  local arr, frames = frame_ptrs(input)
  b2f.check(b2f.lib.b2f_costvol_forward(arr, frames, B, N, h, w, self.win, self.fwd and 1 or 0,
                                        self.output:data(), 0, b2f.stream()))
  return self.output
end

function CostVolMulti:updateGradInput(input, gradOutput)
  local frames = #input
  local ref = input[1]
  local B, N, h, w = ref:size(1), ref:size(2), ref:size(3), ref:size(4)
  for f = 1, frames do
    self.gradInput[f] = self.gradInput[f] or input[f].new()
    self.gradInput[f]:resizeAs(input[f])
  end
  -- gradOutput is a narrow of the 162-channel JoinTable gradient (pwc.lua:267): only the batch
  -- stride differs from a contiguous tensor, and the ABI takes it.
  local go = gradOutput
  if go:stride(2) ~= h * w or go:stride(3) ~= w or go:stride(4) ~= 1 then go = go:contiguous() end
  local arr = frame_ptrs(input)
  local garr = ffi.new('float*[?]', frames)
  for f = 1, frames do garr[f-1] = self.gradInput[f]:data() end
  b2f.check(b2f.lib.b2f_costvol_backward(arr, frames, B, N, h, w, self.win, self.fwd and 1 or 0,
                                         go:data(), go:stride(1), garr, b2f.stream()))
  return self.gradInput
end

function CostVolMulti:clearState()
  return parent.clearState(self)
end

function CostVolMulti:__tostring__()
  return torch.type(self) .. string.format('window size = %d', self.win)
end
