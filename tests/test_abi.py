"""CPU-only checks of the drop-in boundary: the shared library loads without a GPU or driver,
exports exactly the symbols include/b2f.h declares, the ctypes binding lists all of them, and the
host mirror refuses to run without CUDA (no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2f.h")


@pytest.fixture(scope="module")
def built():
    from back2future_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2f_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("b2f_costvol_forward", "b2f_costvol_backward", "b2f_warp_bhwd_forward", "b2f_warp_bhwd_backward",
              "b2f_ob_criterion", "b2f_smoothness_criterion", "b2f_constvel_criterion", "b2f_occprior_criterion",
              "b2f_last_error", "b2f_abi_version"):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol(built):
    lib = C.CDLL(built.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "libb2f_cuda.so does not export %s" % s
    lib.b2f_abi_version.restype = C.c_int
    assert lib.b2f_abi_version() == 1
    lib.b2f_status_string.restype = C.c_char_p
    assert lib.b2f_status_string(0) == b"ok"
    assert lib.b2f_status_string(-1) == b"invalid argument"


def test_binding_covers_header_exactly(built):
    assert sorted(built.SIGNATURES) == declared_symbols()


def test_no_unexpected_exports_and_no_torch_dependency(built):
    out = subprocess.run(["nm", "-D", "--defined-only", built.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared_symbols()
    needed = subprocess.run(["readelf", "-d", built.LIB_PATH], capture_output=True, text=True).stdout
    libs = re.findall(r"NEEDED.*\[(.*?)\]", needed)
    assert not any("torch" in l or "c10" in l or "libcuda.so" in l for l in libs), libs


def test_struct_layouts_match_header(built):
    assert C.sizeof(built.ObParams) == 11 * 4
    assert C.sizeof(built.SmoothParams) == 6 * 4
    assert C.sizeof(built.PackJob) == 3 * 8 + 4 * 4          # b2f_pack_job: three pointers, four int32
    assert built.PackJob.Cout.offset == 24 and built.PackJob.transpose.offset == 36


def test_pack_job_table_bytes(built):
    """_lib.pack_jobs lays the b2f_pack_job array out as the header declares it (the table is uploaded as raw bytes)."""
    raw = built.pack_jobs([(0x1000, 0x2000, 0x3000, 128, 196, 224, 0), (0x10, 0x20, 0x30, 2, 32, 2, 1)])
    assert len(raw) == 2 * 40
    jobs = (built.PackJob * 2).from_buffer_copy(raw)
    assert (jobs[0].w_packed, jobs[0].w_hi, jobs[0].w_lo) == (0x1000, 0x2000, 0x3000)
    assert (jobs[0].Cout, jobs[0].Cin, jobs[0].K, jobs[0].transpose) == (128, 196, 224, 0)
    assert (jobs[1].Cout, jobs[1].Cin, jobs[1].K, jobs[1].transpose) == (2, 32, 2, 1)


def test_argument_validation_needs_no_gpu(built):
    """Rejected arguments are reported before any CUDA call."""
    lib = built.load()
    assert lib.b2f_costvol_forward(None, 2, 1, 1, 4, 4, 9, 1, None, 0, None) == -1
    assert b"frames is NULL" in lib.b2f_last_error()
    assert lib.b2f_warp_bhwd_forward(None, None, None, 1, 4, 4, 3, 4, 4, None) == -1
    assert lib.b2f_occprior_criterion(None, 1, 2, 4, 4, 1.0, 0, None, None, None, None) == -1
    prev = lib.b2f_debug_costvol_path(1)
    assert lib.b2f_debug_costvol_path(prev) == 1
    assert lib.b2f_launch_count(1) == 0
    assert lib.b2f_conv3x3_tc_pack_from_packed_batch(None, 2, None) == -1
    assert lib.b2f_conv3x3_tc_backward_data_s2(None, None, None, None, None, 0, 1, 32, 4, 4, 16, 8, 8, 0, None) == -1


def test_host_mirror_has_reference_surface_and_no_cpu_fallback(built):
    import torch
    from back2future_b200 import nn as bnn
    m = bnn.CostVolMulti()
    assert (m.win, m.fwd, m.verbose) == (3, True, False)            # CostVolMulti.lua:23-47 defaults
    m = bnn.CostVolMulti(9, False)
    assert (m.win, m.fwd) == (9, False) and len(m.gradInput) == 2
    for name in ("updateOutput", "updateGradInput", "accGradParameters", "forward", "backward", "clearState"):
        assert callable(getattr(m, name)) and callable(getattr(bnn.BilinearSamplerBHWD(), name))
    c = bnn.OBGCCriterion()
    assert (c.sizeAverage, c.gradCheck, c.penalty_out, c.alpha, c.beta, c.gamma, c.F, c.pwc_flow_scaling,
            c.past_flow) == (True, False, 1.0, 1.0, 1.0, 1.0, 3, 1, False)
    s = bnn.SecondOrderSmoothnessCriterion()
    assert s.cs == 20 and s.sizeAverage is True and callable(s.clear)
    assert bnn.L1Penalty(0.38).alpha == 0.5                          # L1_function.lua:17
    x = torch.zeros(1, 2, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bnn.CostVolMulti(9).forward([x, x])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bnn.BilinearSamplerBHWD().forward([torch.zeros(1, 4, 4, 2), torch.zeros(1, 4, 4, 2)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bnn.SmoothnessCriterion().forward(x, torch.zeros(1, 3, 4, 4))


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "back2future_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), "%s mentions the oracle" % f
