-- nn.WarpingUnit(flow_scale): the reference's `warpingUnit(I, F)` (models/pwc.lua:68-73) and the
-- nn.MulConstant in front of its flow input (:402-408, :441-446) as ONE module over libb2f_cuda.so
-- (b2f_warp_bdhw_forward / _backward; LuaJIT FFI shim, see INTEGRATION.md, "optional: fused warping unit").
--
-- input = {I (B,C,H,W), F (B,2,H,W)} in the network's own BDHW layout, output (B,C,H,W): the four
-- nn.Transpose copies and the scaling pass of the reference graph disappear.  Results equal the chain
--   {I - Transpose, (F - MulConstant(s)) - Transpose} - nn.BilinearSamplerBHWD() - Transpose
-- within fp32 rounding (tests/test_gpu_parity.py::test_warping_unit_fused_matches_the_reference_chain).
local b2f = require 'b2f_ffi'
local Unit, Base = torch.class('nn.WarpingUnit', 'nn.Module')

function Unit:__init(flow_scale)
  Base.__init(self)
  self.flow_scale = flow_scale or 1
  self.gradInput = {}
end

local function check(I, F, gradOutput)
  assert(I:nDimension() == 4 and F:nDimension() == 4)
  assert(I:size(1) == F:size(1) and F:size(2) == 2, 'flow must be (B,2,H,W)')
  assert(I:size(3) == F:size(3) and I:size(4) == F:size(4), 'image / flow size mismatch')
  if gradOutput then assert(gradOutput:isSameSizeAs(I), 'gradOutput / image size mismatch') end
end

function Unit:updateOutput(input)
  local I, F = input[1]:contiguous(), input[2]:contiguous()
  check(I, F)
  self.output:resizeAs(I)
  b2f.check(b2f.lib.b2f_warp_bdhw_forward(I:data(), F:data(), self.flow_scale, self.output:data(),
                                          I:size(1), I:size(2), I:size(3), I:size(4), b2f.stream()))
  return self.output
end

function Unit:updateGradInput(input, gradOutput)
  local I, F, go = input[1]:contiguous(), input[2]:contiguous(), gradOutput:contiguous()
  check(I, F, go)
  local gI = (self.gradInput[1] or I.new()):resizeAs(I)
  local gF = (self.gradInput[2] or I.new()):resizeAs(F)
  local stream = b2f.stream()
  b2f.check(b2f.lib.b2f_zero_async(gI:data(), gI:nElement() * 4, stream))   -- the entry accumulates into gradImg
  b2f.check(b2f.lib.b2f_warp_bdhw_backward(I:data(), F:data(), self.flow_scale, go:data(), gI:data(), gF:data(),
                                           I:size(1), I:size(2), I:size(3), I:size(4), stream))
  self.gradInput = {gI, gF}
  return self.gradInput
end
