"""ctypes binding of libb2f_cuda.so (the C ABI declared in include/b2f.h).

The product path has no CPU fallback: if the shared library is missing or a call is made
without a CUDA device the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2F_LIB_PATH selects another build of the same library (kernel experiments, tools/build_variants.sh)
LIB_PATH = os.environ.get("B2F_LIB_PATH") or os.path.join(_HERE, "libb2f_cuda.so")

B2F_OK = 0
PENALTY_QUADRATIC, PENALTY_L1, PENALTY_LORENTZIAN = 0, 1, 2

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class ObParams(C.Structure):
    """struct b2f_ob_params (include/b2f.h)."""
    _fields_ = [
        ("gradient_terms", C.c_int), ("penalty", C.c_int), ("penalty_eps", C.c_float),
        ("penalty_out", C.c_float), ("alpha", C.c_float), ("beta", C.c_float), ("gamma", C.c_float),
        ("pwc_flow_scaling", C.c_float), ("past_flow", C.c_int), ("grad_check", C.c_int),
        ("size_average", C.c_int),
    ]


class SmoothParams(C.Structure):
    """struct b2f_smooth_params (include/b2f.h)."""
    _fields_ = [
        ("order", C.c_int), ("penalty", C.c_int), ("penalty_eps", C.c_float), ("cs", C.c_float),
        ("size_average", C.c_int), ("alias_weights", C.c_int),
    ]


# name -> (restype, argtypes); must list every symbol include/b2f.h declares
SIGNATURES = {
    "b2f_abi_version": (C.c_int, []),
    "b2f_last_error": (C.c_char_p, []),
    "b2f_status_string": (C.c_char_p, [C.c_int]),
    "b2f_release_scratch": (C.c_int, []),
    "b2f_reserve_scratch": (C.c_int, [C.c_size_t]),
    "b2f_debug_costvol_path": (C.c_int, [C.c_int]),
    "b2f_zero_async": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "b2f_copy2d_async": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "b2f_launch_count": (C.c_int64, [C.c_int]),
    "b2f_warp_bdhw_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p]),
    "b2f_warp_bdhw_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b2f_flo_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "b2f_flo_read_header": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2f_flo_read": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "b2f_conv3x3_packed_floats": (C.c_int64, [C.c_int, C.c_int]),
    "b2f_conv3x3_pack_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b2f_conv3x3_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_void_p]),
    "b2f_avgpool2x2_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b2f_upsample_bilinear2x_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_float,
                                                  C.c_void_p]),
    "b2f_upsample_nearest_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p]),
    "b2f_softmax_channels_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p]),
    "b2f_conv3x3_tc_packed_floats": (C.c_int64, [C.c_int, C.c_int]),
    "b2f_debug_tc_trace": (C.c_int, [C.c_void_p]),
    "b2f_conv3x3_tc_pack_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "b2f_conv3x3_tc_backward_data": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "b2f_conv3x3_tc_pack_from_packed_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "b2f_conv3x3_tc_backward_data_s2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b2f_conv3x3_tc_backward_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                  C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.c_void_p]),
    "b2f_conv3x3_tc_pack_from_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.c_void_p]),
    "b2f_nhwc_split_from_bdhw": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p]),
    "b2f_conv3x3_tc_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_void_p]),
    "b2f_conv3x3_transpose_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "b2f_conv3x3_backward_data": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                            C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_float, C.c_void_p]),
    "b2f_conv3x3_backward_weights": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b2f_zero_insert2x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "b2f_leaky_relu_backward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_float,
                                          C.c_void_p]),
    "b2f_axpy2d": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p]),
    "b2f_upsample_bilinear2x_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                   C.c_int, C.c_void_p]),
    "b2f_upsample_nearest_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_void_p]),
    "b2f_softmax_channels_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_void_p]),
    "b2f_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_int64, C.c_void_p]),
    "b2f_costvol_forward": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "b2f_costvol_backward": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p),
                                       C.c_void_p]),
    "b2f_warp_bhwd_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b2f_warp_bhwd_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b2f_ob_criterion": (C.c_int, [C.POINTER(ObParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_double_p, C.c_void_p]),
    "b2f_smoothness_criterion": (C.c_int, [C.POINTER(SmoothParams), C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                           c_double_p, C.c_void_p]),
    "b2f_constvel_criterion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_double_p,
                                         C.c_void_p]),
    "b2f_occprior_criterion": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                         C.c_int, C.c_void_p, C.c_void_p, c_double_p, C.c_void_p]),
}


class B2FError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libb2f_cuda: status %d: %s" % (status, message))
        self.status = status


_lib = None


def load():
    """Load libb2f_cuda.so and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libb2f_cuda.so is not built (%s); run `make` or `python -c 'import __graft_entry__ as g; "
            "g.build()'`.  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.b2f_abi_version() != 1:
        raise ImportError("libb2f_cuda.so ABI version %d, expected 1" % lib.b2f_abi_version())
    _lib = lib
    return lib


def check(status):
    if status != B2F_OK:
        lib = load()
        msg = lib.b2f_last_error().decode("utf-8", "replace") or lib.b2f_status_string(status).decode()
        raise B2FError(status, msg)


class PackJob(C.Structure):
    """b2f_pack_job (include/b2f.h)."""
    _fields_ = [("w_packed", C.c_void_p), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("Cout", C.c_int32),
                ("Cin", C.c_int32), ("K", C.c_int32), ("transpose", C.c_int32)]


def pack_jobs(jobs):
    """Host bytes of a b2f_pack_job array; jobs = [(w_packed_ptr, hi_ptr, lo_ptr, Cout, Cin, K, transpose), ...]."""
    arr = (PackJob * len(jobs))()
    for i, j in enumerate(jobs):
        arr[i] = PackJob(*j)
    return bytes(arr)


def ptr_array(ptrs):
    """Host array of (device) pointers; None -> NULL."""
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
