-- LuaJIT FFI binding of libb2f_cuda.so (include/b2f.h).  UNTESTED in this repository: no LuaJIT /
-- Torch7 exists in the build image; the same C ABI is exercised from Python ctypes in tests/.
-- The handle lives in a file-local upvalue so that modules stay torch.save-serialisable
-- (no cdata in `self`; SURVEY 8b "Serialisation constraint").
local ffi = require 'ffi'
require 'cutorch'

ffi.cdef[[
typedef struct CUstream_st* b2f_stream_t;
typedef struct b2f_ob_params {
  int gradient_terms; int penalty; float penalty_eps; float penalty_out;
  float alpha, beta, gamma; float pwc_flow_scaling; int past_flow; int grad_check; int size_average;
} b2f_ob_params;
typedef struct b2f_smooth_params {
  int order; int penalty; float penalty_eps; float cs; int size_average; int alias_weights;
} b2f_smooth_params;
int b2f_abi_version(void);
const char* b2f_last_error(void);
int b2f_zero_async(void* ptr, size_t bytes, b2f_stream_t stream);
int b2f_release_scratch(void);
int b2f_reserve_scratch(size_t bytes);
int b2f_costvol_forward(const float* const* frames, int F, int B, int C, int H, int W, int win, int fwd,
                        float* out, int64_t out_batch_stride, b2f_stream_t stream);
int b2f_costvol_backward(const float* const* frames, int F, int B, int C, int H, int W, int win, int fwd,
                         const float* gradOut, int64_t gradOut_batch_stride, float* const* gradFrames,
                         b2f_stream_t stream);
int b2f_warp_bhwd_forward(const float* img, const float* grid, float* out, int B, int H, int W, int C,
                          int Hg, int Wg, b2f_stream_t stream);
int b2f_warp_bhwd_backward(const float* img, const float* grid, const float* gradOut, float* gradImg,
                           float* gradGrid, int B, int H, int W, int C, int Hg, int Wg, b2f_stream_t stream);
int b2f_warp_bdhw_forward(const float* img, const float* flow, float flow_scale, float* out, int B, int C,
                          int H, int W, b2f_stream_t stream);
int b2f_warp_bdhw_backward(const float* img, const float* flow, float flow_scale, const float* gradOut,
                           float* gradImg, float* gradFlow, int B, int C, int H, int W, b2f_stream_t stream);
int b2f_ob_criterion(const b2f_ob_params* prm, const float* flow, const float* bflow, const float* occ,
                     const float* warp_past, const float* warp_future, const float* target,
                     int B, int C, int h, int w, float* grad_occ, float* grad_warp_past,
                     float* grad_warp_future, double* loss_dev, double* loss_host, b2f_stream_t stream);
int b2f_smoothness_criterion(const b2f_smooth_params* prm, const float* input, const float* target,
                             int B, int Cin, int Ct, int h, int w, float* grad, double* loss_dev,
                             double* loss_host, b2f_stream_t stream);
int b2f_constvel_criterion(const float* f, const float* b, int B, int C, int h, int w, int size_average,
                           float* grad_f, float* grad_b, double* loss_dev, double* loss_host,
                           b2f_stream_t stream);
int b2f_occprior_criterion(const float* occ, int B, int C, int h, int w, float penalty, int size_average,
                           float* grad, double* loss_dev, double* loss_host, b2f_stream_t stream);
/* cutorch: current stream of the calling thread's device */
typedef struct THCState THCState;
b2f_stream_t THCState_getCurrentStream(THCState* state);
]]

local lib = ffi.load(os.getenv('B2F_CUDA_LIB') or 'b2f_cuda')
assert(lib.b2f_abi_version() == 1, 'libb2f_cuda ABI mismatch')

local M = { lib = lib, ffi = ffi }

function M.stream()
  return ffi.C.THCState_getCurrentStream(cutorch.getState())
end

-- turn a status into a Lua error, like THError("aborting") in the reference (.cu:152-156)
function M.check(status)
  if status ~= 0 then error('libb2f_cuda: ' .. ffi.string(lib.b2f_last_error()), 2) end
end

function M.ptr(t) return t and t:data() or nil end   -- cutorch FFI: float* of a CudaTensor

local penalty_kind = { QuadraticPenalty = 0, L1Penalty = 1, LorentzianPenalty = 2 }
function M.penalty(p)
  local kind = penalty_kind[torch.type(p)]
  assert(kind, 'unsupported penalty ' .. torch.type(p))
  return kind, (kind == 2) and p.eps or 0
end

return M
