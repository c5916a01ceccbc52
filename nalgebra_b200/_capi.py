"""ctypes binding of libnalgebra_b200.so (the C ABI declared in include/nalgebra_b200.h).

The library is the product; this module only loads it and declares argument types.  There is no
CPU fallback anywhere: if the shared library is missing, or no sm_100 device is present, calls
raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NAB_LIB") or os.path.join(_HERE, "libnalgebra_b200.so")

NA_OK, NA_NOT_PD, NA_SINGULAR = 0, 1, 2
NA_EINVAL, NA_ECUDA, NA_ENOMEM, NA_ENCCL = -1, -2, -3, -4

_sz, _pd, _dbl, _flt, _int, _p, _u64 = C.c_size_t, C.c_ssize_t, C.c_double, C.c_float, C.c_int, C.c_void_p, C.c_uint64

# name -> (restype, argtypes); mirrors include/nalgebra_b200.h one to one.
_GEMM64 = [_sz, _sz, _sz, _dbl, _p, _pd, _pd, _p, _pd, _pd, _dbl, _p, _pd, _pd]
_GEMM32 = [_sz, _sz, _sz, _flt, _p, _pd, _pd, _p, _pd, _pd, _flt, _p, _pd, _pd]
SIGNATURES = {
    "na_init": (_int, [_int]),
    "na_shutdown": (_int, []),
    "na_last_error": (C.c_char_p, []),
    "na_version": (C.c_char_p, []),
    "na_kernel_launches": (_u64, []),
    "na_dev_malloc": (_int, [C.POINTER(_p), _sz]),
    "na_dev_free": (_int, [_p]),
    "na_host_alloc_pinned": (_int, [C.POINTER(_p), _sz]),
    "na_host_free_pinned": (_int, [_p]),
    "na_memcpy_h2d": (_int, [_p, _p, _sz]),
    "na_memcpy_d2h": (_int, [_p, _p, _sz]),
    "na_dev_synchronize": (_int, []),
    "na_ipc_get_handle": (_int, [_p, _p]),
    "na_ipc_open_handle": (_int, [_p, C.POINTER(_p)]),
    "na_ipc_close_handle": (_int, [_p]),
    "na_memcpy_peer_async": (_int, [_p, _p, _sz, _p]),
    "na_fill_uniform_dev": (_int, [_p, _sz, _sz, _sz, _u64, _p]),
    "na_fill_uniform_block_dev": (_int, [_p, _sz, _sz, _sz, _u64, _sz, _sz, _sz, _p]),
    "na_dgemm": (_int, _GEMM64),
    "na_sgemm": (_int, _GEMM32),
    "na_dgemm_dev": (_int, _GEMM64 + [_p]),
    "na_sgemm_dev": (_int, _GEMM32 + [_p]),
    "na_cholesky_f64": (_int, [_sz, _p, _sz, _int, _dbl, _p]),
    "na_cholesky_f64_dev": (_int, [_sz, _p, _sz, _int, _dbl, _p, _p]),
    "na_cholesky_solve_f64": (_int, [_sz, _p, _sz, _p, _sz, _sz]),
    "na_cholesky_solve_f64_dev": (_int, [_sz, _p, _sz, _p, _sz, _sz, _p]),
    "na_lu_f64": (_int, [_sz, _sz, _p, _sz, _p, _p]),
    "na_lu_f64_dev": (_int, [_sz, _sz, _p, _sz, _p, _p, _p]),
    "na_lu_solve_f64": (_int, [_sz, _p, _sz, _p, _sz, _p, _sz, _sz]),
    "na_lu_solve_f64_dev": (_int, [_sz, _p, _sz, _p, _sz, _p, _sz, _sz, _p]),
    "na_qr_f64": (_int, [_sz, _sz, _p, _sz, _p]),
    "na_qr_f64_dev": (_int, [_sz, _sz, _p, _sz, _p, _p]),
    "na_qr_q_f64": (_int, [_sz, _sz, _p, _sz, _p, _p, _sz]),
    "na_qr_q_f64_dev": (_int, [_sz, _sz, _p, _sz, _p, _p, _sz, _p]),
    "na_qr_q_tr_mul_f64": (_int, [_sz, _sz, _p, _sz, _p, _p, _sz, _sz]),
    "na_qr_q_tr_mul_f64_dev": (_int, [_sz, _sz, _p, _sz, _p, _p, _sz, _sz, _p]),
    "na_qr_solve_f64": (_int, [_sz, _p, _sz, _p, _p, _sz, _sz]),
    "na_full_piv_lu_f64": (_int, [_sz, _sz, _p, _sz, _p, _p, _p, _p]),
    "na_full_piv_lu_f64_dev": (_int, [_sz, _sz, _p, _sz, _p, _p, _p, _p, _p]),
    "na_col_piv_qr_f64": (_int, [_sz, _sz, _p, _sz, _p, _p, _p]),
    "na_col_piv_qr_f64_dev": (_int, [_sz, _sz, _p, _sz, _p, _p, _p, _p]),
    "na_hessenberg_f64": (_int, [_sz, _p, _sz, _p]),
    "na_hessenberg_f64_dev": (_int, [_sz, _p, _sz, _p, _p]),
    "na_symmetric_tridiagonal_f64": (_int, [_sz, _p, _sz, _p]),
    "na_symmetric_tridiagonal_f64_dev": (_int, [_sz, _p, _sz, _p, _p]),
    "na_bidiagonal_f64": (_int, [_sz, _sz, _p, _sz, _p, _p]),
    "na_bidiagonal_f64_dev": (_int, [_sz, _sz, _p, _sz, _p, _p, _p]),
    "na_set_gemm_sm_limit": (_int, [_int]),
    "na_trsm_f64_dev": (_int, [_int, _int, _int, _int, _sz, _sz, _p, _sz, _p, _sz, _p]),
    "na_permute_rows_f64_dev": (_int, [_sz, _p, _sz, _sz, _p, _sz, _int, _p]),
    "na_dgemm_lower_dev": (_int, _GEMM64[:-2] + [_sz, _p]),
    "na_fill_spd_block_dev": (_int, [_p, _sz, _sz, _sz, _u64, _sz, _sz, _sz, _p]),
    "na_dsyrk_lower": (_int, [_sz, _sz, _dbl, _p, _pd, _pd, _dbl, _p, _sz]),
    "na_dgemv": (_int, [_sz, _sz, _dbl, _p, _pd, _pd, _p, _pd, _dbl, _p, _pd]),
    "na_dgemv_dev": (_int, [_sz, _sz, _dbl, _p, _pd, _pd, _p, _pd, _dbl, _p, _pd, _p]),
    "na_daxcpy_dev": (_int, [_sz, _dbl, _p, _pd, _dbl, _dbl, _p, _pd, _p]),
    "na_cholesky_f64_dev_async": (_int, [_sz, _p, _sz, _int, _dbl, _p, _sz, _p]),
    "na_lu_f64_dev_async": (_int, [_sz, _sz, _p, _sz, _p, _p]),
    "na_apply_ipiv_f64_dev": (_int, [_sz, _p, _sz, _sz, _p, _sz, _sz, _p]),
    "na_set_tuning": (_int, [C.c_char_p, C.c_long]),
    # Fortran-ABI LAPACK facade: every argument by pointer
    "dpotrf_": (None, [_p] * 5), "dpotrs_": (None, [_p] * 8), "dpotri_": (None, [_p] * 5),
    "dgetrf_": (None, [_p] * 6), "dlaswp_": (None, [_p] * 7), "dgetrs_": (None, [_p] * 9), "dgetri_": (None, [_p] * 7),
    "dgeqrf_": (None, [_p] * 8), "dormqr_": (None, [_p] * 13), "dorgqr_": (None, [_p] * 9), "dtrtrs_": (None, [_p] * 10),
    "na_tri_solve_f64": (_int, [_int, _int, _int, _sz, _p, _sz, _p, _sz, _sz]),
    "na_tri_solve_f64_dev": (_int, [_int, _int, _int, _sz, _p, _sz, _p, _sz, _sz, _p]),
}


class NalgebraB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libnalgebra_b200 status {status}: {message}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """Loads the shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m nalgebra_b200.build` "
                "(nalgebra_b200 has no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name, None)
            if fn is None:
                continue  # test_capi_symbols reports missing exports
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def last_error() -> str:
    return lib().na_last_error().decode(errors="replace")


def check(status: int) -> int:
    """Raises on negative (error) statuses; returns non-negative ones (NA_OK / NA_NOT_PD / NA_SINGULAR)."""
    if status < 0:
        raise NalgebraB200Error(status, last_error())
    return status
