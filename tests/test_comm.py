"""libb2f_comm.so (include/b2f_comm.h): exports, binding table, and -- on a GPU -- a world-1 communicator."""
import ctypes as C
import os
import re
import subprocess

import pytest

from back2future_b200 import comm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_comm_library_exports_what_the_header_declares():
    hdr = open(os.path.join(ROOT, "include", "b2f_comm.h")).read()
    declared = set(re.findall(r"B2F_COMM_API\s+[\w\s\*]+?\b(b2f_comm_\w+)\s*\(", hdr))
    assert declared == set(comm.SIGNATURES), declared ^ set(comm.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", comm.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported
    lib = comm.load()
    assert lib.b2f_comm_abi_version() == 1
    # neither NCCL nor torch is a link-time dependency (NCCL is dlopen'ed on first use)
    needed = subprocess.run(["readelf", "-d", comm.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "nccl" not in needed and "torch" not in needed
    lua = open(os.path.join(ROOT, "lua", "b2f_comm.lua")).read()
    for name in declared:
        assert name in lua, name


def test_comm_argument_checks():
    lib = comm.load()
    assert lib.b2f_comm_unique_id(None) != 0
    h = C.c_void_p()
    assert lib.b2f_comm_init(C.byref(h), None, 1, 0) != 0
    assert lib.b2f_comm_init(C.byref(h), b"\0" * 128, 2, 5) != 0
    assert b"world" in lib.b2f_comm_last_error()
    assert lib.b2f_comm_allreduce_sum_f32(None, None, 4, None) != 0
    assert lib.b2f_comm_destroy(None) == 0
    with pytest.raises(ValueError):
        comm.Communicator(b"short", 1, 0)


@pytest.mark.gpu
def test_comm_world_one_on_the_gpu():
    import torch
    c = comm.Communicator(comm.unique_id(), 1, 0)
    t = torch.arange(10, device="cuda", dtype=torch.float32)
    c.allreduce_sum(t)
    c.allreduce_sum(t, 2, 5)
    torch.cuda.synchronize()
    assert t.tolist() == list(range(10))
    c.destroy()
