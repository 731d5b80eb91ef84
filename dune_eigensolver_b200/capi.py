"""ctypes binding of the C ABI (include/dune_eigensolver_b200.h) -- thin, no logic.

The shared library is built in-tree by dune_eigensolver_b200.build. There is no CPU fallback: if the library
is missing this module raises, and every compute call fails with DE_ERR_CUDA on a machine without a GPU.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

DE_OK, DE_ERR_INVALID, DE_ERR_ALLOC, DE_ERR_CUDA, DE_ERR_NCCL, DE_ERR_SINGULAR, DE_ERR_UNSUPPORTED = range(7)
DE_MAX_COLS = 64
DE_SPMM_AUTO, DE_SPMM_CSR, DE_SPMM_BRB = 0, 1, 2

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_lp = C.POINTER(C.c_long)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

# name -> (argtypes)   ; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "de_version": [],
    "de_last_error_string": [_vp],
    "de_context_create": [C.c_int, _vp, _vpp],
    "de_context_destroy": [_vp],
    "de_context_synchronize": [_vp],
    "de_context_launch_count": [_vp, _i64p],
    "de_context_set_profiling": [_vp, C.c_int],
    "de_context_set_option": [_vp, C.c_char_p, C.c_int64],
    "de_context_profile": [_vp, C.c_int, _dp, _i64p, C.c_int],
    "de_comm_unique_id": [_vp],
    "de_context_init_comm": [_vp, C.c_int, C.c_int, _vp],
    "de_context_rank": [_vp, _ip, _ip],
    "de_context_peer_window_create": [_vp, C.c_int64, _vp],
    "de_context_peer_window_open": [_vp, _vp],
    "de_context_peer_ready": [_vp, _ip],
    "de_matrix_set_peer_deposit": [_vp, _i64p, C.c_int64],
    "de_matrix_create_csr": [_vp, C.c_int64, C.c_int64, _i64p, _i64p, _dp, _vpp],
    "de_matrix_create_bcsr": [_vp, C.c_int64, C.c_int64, C.c_int, _i64p, _i64p, _dp, _vpp],
    "de_matrix_create_distributed": [_vp, C.c_int64, C.c_int64, C.c_int64, _i64p, _i64p, _dp, C.c_int, _ip, _i64p,
                                     _i64p, _i64p, _vpp],
    "de_halo_plan_peers": [C.c_int, C.c_int, _i64p, _i64p, _i64p, C.c_int64, _ip, _ip, _i64p, _i64p, _i64p, _i64p, _i64p, _ip],
    "de_matrix_create_rowblock": [_vp, C.c_int64, C.c_int64, _i64p, _i64p, _dp, _i64p, _vp, _vp, _vpp],
    "de_multi_create": [_ip, C.c_int, C.c_int64, _vpp],
    "de_multi_destroy": [_vp],
    "de_multi_size": [_vp, _ip],
    "de_multi_set_timeout": [_vp, C.c_double],
    "de_multi_context": [_vp, C.c_int, _vpp],
    "de_multi_last_error": [_vp],
    "de_multi_launch_count": [_vp, _i64p],
    "de_multi_standard_largest": [_vp, C.c_int64, C.c_int64, _i64p, _i64p, _dp, C.c_int64, C.c_double, C.c_double,
                                  C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip],
    "de_multi_standard_lobpcg": [_vp, C.c_int64, C.c_int64, _i64p, _i64p, _dp, C.c_int64, C.c_double, C.c_int, C.c_int,
                                 _dp, _dp, _dp, C.c_int, _ip],
    "de_multi_generalized_lobpcg": [_vp, C.c_int64, C.c_int64, _i64p, _i64p, _dp, C.c_int64, _i64p, _i64p, _dp,
                                    C.c_int64, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip],
    "de_matrix_destroy": [_vp],
    "de_matrix_rows": [_vp, _i64p, _i64p],
    "de_matrix_set_spmm_format": [_vp, C.c_int],
    "de_matrix_spmm_info": [_vp, _ip, _i64p, _i64p, _i64p, _i64p, _ip],
    "de_matrix_brb_selfcheck": [_vp, C.c_int64, C.c_int64, _i64p, _i64p, _dp, _i64p],
    "de_brb_format_check": [C.c_int64, C.c_int64, C.c_int64, _i64p, _i64p, _dp, C.c_int, _i64p, _dp],
    "de_halo_plan_local": [C.c_int64, _i64p, _i64p, C.c_int, C.c_int, _i64p, _i64p, _i64p, _i64p, _i64p],
    "de_mv_create": [_vp, C.c_int64, C.c_int, _vpp],
    "de_mv_destroy": [_vp],
    "de_mv_shape": [_vp, _i64p, _ip],
    "de_mv_upload_panel8": [_vp, _dp],
    "de_mv_download_panel8": [_vp, _dp],
    "de_mv_upload_rowmajor": [_vp, _dp],
    "de_mv_download_rowmajor": [_vp, _dp],
    "de_mv_copy": [_vp, _vp],
    "de_mv_device_ptr": [_vp, _vpp],
    "de_spmm": [_vp, _vp, _vp],
    "de_spmm_diag_dot": [_vp, _vp, _vp, _dp],
    "de_spmm_gram": [_vp, _vp, _vp, _dp, _dp],
    "de_diag_dot": [_dp, _vp, _vp],
    "de_gram": [_dp, _vp, _vp],
    "de_block_update": [_vp, _dp],
    "de_block_project": [_vp, C.c_int, C.c_int, C.c_int, _dp],
    "de_orthonormalize": [_vp],
    "de_b_orthonormalize": [_vp, _vp, _vp, _dp],
    "de_factor_upload": [_vp, C.c_int64, _lp, _lp, _dp, _lp, _lp, _dp, _lp, _lp, _dp, C.c_long, _vpp],
    "de_factor_destroy": [_vp],
    "de_factor_apply": [_vp, _vp, _vp],
    "de_factor_info": [_vp, _i64p, _i64p, _i64p, _ip, _ip],
    "de_standard_largest": [_vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip],
    "de_standard_largest_mv": [_vp, _vp, C.c_double, C.c_double, C.c_int, _vp, _dp, C.c_int, _ip],
    "de_standard_inverse_mv": [_vp, _vp, _vp, C.c_double, C.c_double, C.c_int, _vp, _dp, C.c_int, _ip],
    "de_standard_inverse": [_vp, _vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip],
    "de_generalized_inverse": [_vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int,
                               _ip, _dp],
    "de_standard_lobpcg": [_vp, _vp, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip],
    "de_generalized_lobpcg": [_vp, _vp, _vp, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip],
    "de_lobpcg_mv": [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _dp, _dp, C.c_int, _ip, _ip, _ip],
    "de_block_lincomb": [_vp, _vp, C.c_int, _vpp, _dp],
    "de_host_sym_eig": [C.c_int, _dp, _dp, _dp],
    "de_host_sym_gen_eig": [C.c_int, _dp, _dp, _dp, _dp, _dp],
    "de_start_block": [C.c_int64, C.c_int, C.c_uint, _dp],
    "de_host_factorize": [C.c_int64, _i64p, _i64p, _dp, C.c_int, C.c_int, _vpp],
    "de_host_factor_arrays": [_vp, _i64p, _i64p, _i64p, C.POINTER(_lp), C.POINTER(_lp), C.POINTER(_dp),
                              C.POINTER(_lp), C.POINTER(_lp), C.POINTER(_dp), C.POINTER(_lp), C.POINTER(_lp),
                              C.POINTER(_dp), _lp],
    "de_host_factor_destroy": [_vp],
    "de_host_factorize_spd": [C.c_int64, _i64p, _i64p, _dp, C.c_int, C.c_int, _vpp],
    "de_host_factor_info": [_vp, _ip, _i64p, _i64p, _i64p, _dp, _dp],
    "de_factor_upload_host": [_vp, _vp, _vpp],
}
_RESTYPES = {"de_last_error_string": C.c_char_p, "de_multi_last_error": C.c_char_p}
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)

_lib = None


class DeError(RuntimeError):
    """A C-ABI call failed. `.status` is the de_status code. DE_ERR_INVALID mirrors the reference's
    std::invalid_argument (same message text where the reference has one)."""

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def library_path():
    return _build.LIB


def lib():
    """Load (building first if needed) the shared library. Fails loudly if it cannot be had."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            path = _build.build_library()
        L = C.CDLL(path)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = L
    return _lib


def check(status, ctx=None):
    if status != DE_OK:
        msg = lib().de_last_error_string(ctx)
        raise DeError(status, (msg or b"").decode() or "de_status %d" % status)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def dptr(a):
    return a.ctypes.data_as(_dp)


def i64ptr(a):
    return a.ctypes.data_as(_i64p)


def lptr(a):
    return a.ctypes.data_as(_lp)


def i32ptr(a):
    return a.ctypes.data_as(_ip)
