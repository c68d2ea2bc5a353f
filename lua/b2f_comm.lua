-- LuaJIT FFI binding of libb2f_comm.so (include/b2f_comm.h): the gradient all-reduce of the training path.
-- UNTESTED here (no LuaJIT / Torch7 in the build image); the same ABI is exercised from Python (tests/test_comm.py).
-- Replaces util.lua:27-48 (DataParallelTable with usenccl) + train.lua:494-496 (syncParameters) when main.lua is
-- launched as one process per GPU: after model:backward, call comm.allreduce(gradParams) and run optim.adam on
-- every rank.
local ffi = require 'ffi'
require 'cutorch'

ffi.cdef[[
typedef struct b2f_comm* b2f_comm_t;
int b2f_comm_abi_version(void);
const char* b2f_comm_last_error(void);
int b2f_comm_unique_id(void* id_out);
int b2f_comm_init(b2f_comm_t* comm, const void* id, int world, int rank);
int b2f_comm_world(b2f_comm_t comm, int* world, int* rank);
int b2f_comm_allreduce_sum_f32(b2f_comm_t comm, float* buf, size_t count, void* stream);
int b2f_comm_allreduce_sum_f64(b2f_comm_t comm, double* buf, size_t count, void* stream);
int b2f_comm_destroy(b2f_comm_t comm);
]]

local lib = ffi.load(os.getenv('B2F_COMM_LIB_PATH') or 'b2f_comm')
assert(lib.b2f_comm_abi_version() == 1, 'libb2f_comm ABI mismatch')

local function check(rc)
  if rc ~= 0 then error('libb2f_comm: ' .. ffi.string(lib.b2f_comm_last_error())) end
end

local M = {}

function M.uniqueId()                 -- rank 0; ship the 128-byte string to the other ranks
  local id = ffi.new('char[128]')
  check(lib.b2f_comm_unique_id(id))
  return ffi.string(id, 128)
end

function M.init(id, world, rank)      -- on the process's current device (cutorch.setDevice first)
  local h = ffi.new('b2f_comm_t[1]')
  check(lib.b2f_comm_init(h, id, world, rank))
  return h[0]
end

function M.allreduce(comm, flatGrad)  -- flatGrad: the contiguous CudaTensor getParameters() returned
  assert(flatGrad:isContiguous(), 'flattened gradient must be contiguous')
  local stream = require('b2f_ffi').stream()     -- cutorch's current stream (lua/b2f_ffi.lua)
  check(lib.b2f_comm_allreduce_sum_f32(comm, flatGrad:data(), flatGrad:nElement(), stream))
end

function M.destroy(comm) check(lib.b2f_comm_destroy(comm)) end

return M
