"""ctypes binding of libb2f_comm.so (include/b2f_comm.h): the training path's gradient all-reduce over NCCL, behind
the C ABI a LuaJIT host would bind (lua/b2f_comm.lua) -- not torch.distributed's collective.

`Communicator.from_env()` does the rendezvous for the Python mirror: rank 0 draws the NCCL unique id and publishes
it through torch.distributed's key-value store (plumbing; the reference's host would use its `threads` channel).
Replaces util.lua:27-48 (`nn.DataParallelTable(1, true, true)`, usenccl) + train.lua:494-496 (`syncParameters`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2F_COMM_LIB_PATH") or os.path.join(_HERE, "libb2f_comm.so")
ID_BYTES = 128

SIGNATURES = {
    "b2f_comm_abi_version": (C.c_int, []),
    "b2f_comm_last_error": (C.c_char_p, []),
    "b2f_comm_unique_id": (C.c_int, [C.c_void_p]),
    "b2f_comm_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int]),
    "b2f_comm_world": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2f_comm_allreduce_sum_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b2f_comm_allreduce_sum_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b2f_comm_destroy": (C.c_int, [C.c_void_p]),
}

_lib = None


class CommError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libb2f_comm.so is not built (%s); run `make`" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.b2f_comm_abi_version() != 1:
            raise ImportError("libb2f_comm.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise CommError("libb2f_comm: status %d: %s" % (rc, load().b2f_comm_last_error().decode("utf-8", "replace")))


def unique_id():
    buf = C.create_string_buffer(ID_BYTES)
    check(load().b2f_comm_unique_id(buf))
    return buf.raw


class Communicator:
    """One NCCL communicator on the current CUDA device."""

    def __init__(self, uid: bytes, world: int, rank: int):
        if len(uid) != ID_BYTES:
            raise ValueError("unique id must be %d bytes" % ID_BYTES)
        self._h = C.c_void_p()
        self.world, self.rank = int(world), int(rank)
        check(load().b2f_comm_init(C.byref(self._h), uid, self.world, self.rank))

    @classmethod
    def from_env(cls, store_key="b2f_comm_uid"):
        """World / rank from torch.distributed (already initialised by the launcher) or a single-rank communicator."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return cls(unique_id(), 1, 0)
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(box[0], world, rank)

    def allreduce_sum(self, t, lo=0, hi=None, stream=None):
        """In-place sum of t.view(-1)[lo:hi] (a contiguous CUDA float32 / float64 tensor) on `stream` (default: the
        current torch stream)."""
        import torch
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("allreduce_sum: contiguous CUDA tensor expected")
        n = t.numel()
        hi = n if hi is None else hi
        if not (0 <= lo <= hi <= n):
            raise ValueError("allreduce_sum: bad range [%d, %d) of %d" % (lo, hi, n))
        st = C.c_void_p((stream or torch.cuda.current_stream(t.device)).cuda_stream)
        ptr = C.c_void_p(t.data_ptr() + lo * t.element_size())
        if t.dtype == torch.float32:
            check(load().b2f_comm_allreduce_sum_f32(self._h, ptr, hi - lo, st))
        elif t.dtype == torch.float64:
            check(load().b2f_comm_allreduce_sum_f64(self._h, ptr, hi - lo, st))
        else:
            raise TypeError("allreduce_sum: float32 / float64 only")

    def destroy(self):
        if self._h:
            check(load().b2f_comm_destroy(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
