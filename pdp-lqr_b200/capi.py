"""ctypes binding of the C ABI declared in include/pdplqr.h (libpdplqr.so).  There is no CPU fallback: if the
library is missing it is (re)built with nvcc, and if that fails the import error propagates."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_ORDER, ERR_NOT_PD = 0, -1, -2, -3, -4, -5
CONDENSED_LU, CONDENSED_CHOLESKY = 0, 1
OPT_AFFINE_CACHE = 1
OPT_INTERIOR_SHARD = 2

_dp = C.c_void_p   # double* (host or device address)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/pdplqr.h declares (checked by tests/test_capi.py)
SIGNATURES = {
    "pdplqr_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, _ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pdplqr_destroy": (C.c_int, [C.c_void_p]),
    "pdplqr_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdplqr_set_model": (C.c_int, [C.c_void_p] + [_dp] * 7),
    "pdplqr_set_model_device": (C.c_int, [C.c_void_p] + [_dp] * 7),
    "pdplqr_update_problem_data": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double]),
    "pdplqr_backward": (C.c_int, [C.c_void_p, _dp]),
    "pdplqr_backward_without_factorization": (C.c_int, [C.c_void_p, _dp]),
    "pdplqr_forward": (C.c_int, [C.c_void_p, _dp, _dp]),
    "pdplqr_solve": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp]),
    "pdplqr_update_problem_data_device": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double]),
    "pdplqr_backward_device": (C.c_int, [C.c_void_p, _dp]),
    "pdplqr_backward_without_factorization_device": (C.c_int, [C.c_void_p, _dp]),
    "pdplqr_forward_device": (C.c_int, [C.c_void_p, _dp, _dp]),
    "pdplqr_solve_device": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp]),
    "pdplqr_synchronize": (C.c_int, [C.c_void_p]),
    "pdplqr_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "pdplqr_summary_doubles": (C.c_int, [C.c_void_p]),
    "pdplqr_get_root_summary_device": (C.c_int, [C.c_void_p, _dp]),
    "pdplqr_set_root_boundary_device": (C.c_int, [C.c_void_p, _dp, _dp]),
    "pdplqr_coupler_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pdplqr_coupler_solve_device": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp]),
    "pdplqr_sharded_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, _ip, C.c_int, _ip, C.c_int, C.c_int]),
    "pdplqr_sharded_set_model": (C.c_int, [C.c_void_p] + [_dp] * 7),
    "pdplqr_sharded_solve": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp]),
    "pdplqr_sharded_num_devices": (C.c_int, [C.c_void_p]),
    "pdplqr_sharded_last_error": (C.c_char_p, [C.c_void_p]),
    "pdplqr_sharded_destroy": (C.c_int, [C.c_void_p]),
    "pdplqr_admm_set_cones": (C.c_int, [C.c_void_p, C.c_int, _ip, _ip, _ip, _ip, _dp, _dp]),
    "pdplqr_admm_solve": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_double,
                                    C.c_double, C.c_int, _ip, _dp]),
    "pdplqr_admm_solve_device": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int,
                                           C.c_double, C.c_double, C.c_int, _ip, _dp]),
    "pdplqr_admm_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]),
    "pdplqr_admm_stats": (C.c_int, [C.c_void_p, _ip, _ip]),
    "pdplqr_num_segments": (C.c_int, [C.c_void_p]),
    "pdplqr_get_partition": (C.c_int, [C.c_void_p, _ip, _ip]),
    "pdplqr_get_gains": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "pdplqr_get_interface": (C.c_int, [C.c_void_p, _dp, _dp]),
    "pdplqr_get_summaries": (C.c_int, [C.c_void_p] + [_dp] * 5),
    "pdplqr_get_costates": (C.c_int, [C.c_void_p, _dp, _dp]),
    "pdplqr_get_costates_device": (C.c_int, [C.c_void_p, _dp, _dp]),
    "pdplqr_last_status": (C.c_int, [C.c_void_p, _ip]),
    "pdplqr_last_error": (C.c_char_p, [C.c_void_p]),
    "pdplqr_launch_count": (C.c_longlong, [C.c_void_p]),
    "pdplqr_debug_check_guards": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    "pdplqr_record_doubles": (C.c_int, [C.c_void_p, _ip, _ip]),
    "pdplqr_wave_size": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "pdplqr_version": (C.c_int, []),
}

_LIB = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load libpdplqr.so (building it in-tree first if it is missing or stale and nvcc is available)."""
    global _LIB
    if _LIB is None:
        if _build.stale() and os.path.exists(_build.NVCC):
            _build.build()
        L = C.CDLL(_build.LIB)   # raises OSError if absent: the product path fails loudly, no fallback
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB
