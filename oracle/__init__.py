"""CPU parity oracle for nalgebra's dense hot path -- TEST INFRASTRUCTURE ONLY.

ctypes binding over ``oracle/libnalgebra_oracle.so`` (built from ``nalgebra_oracle.c`` by
``oracle/Makefile`` / ``__graft_entry__.build()``).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module, and only as
the checker or the CPU baseline.  The product (``nalgebra_b200``) never imports it.

Every function mirrors one reference routine; see the C header for file:line citations.
All matrices are numpy float64 arrays in Fortran (column-major) order.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnalgebra_oracle.so")

_sz, _pd, _dbl, _int = C.c_size_t, C.c_ssize_t, C.c_double, C.c_int
_p = C.c_void_p


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``-O2 -ffp-contract=off``)."""
    src = os.path.join(_HERE, "nalgebra_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "CC=gcc"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        gemm_sig = [_sz, _sz, _sz, _dbl, _p, _pd, _pd, _p, _pd, _pd, _dbl, _p, _pd, _pd]
        for name in ("na_oracle_gemm_f64", "na_oracle_gemm_fallback_f64", "na_oracle_gemm_tr_f64"):
            getattr(_lib, name).argtypes = gemm_sig
            getattr(_lib, name).restype = None
        _lib.na_oracle_dgemm_mm.argtypes = gemm_sig + [_int]
        _lib.na_oracle_dgemm_mm.restype = None
        _lib.na_oracle_gemm_f32.argtypes = [_sz, _sz, _sz, C.c_float, _p, _pd, _pd, _p, _pd, _pd, C.c_float, _p, _pd, _pd]
        _lib.na_oracle_gemm_f32.restype = None
        _lib.na_oracle_rand01.argtypes = [C.c_uint64, C.c_uint64]
        _lib.na_oracle_rand01.restype = _dbl
        _lib.na_oracle_fill_uniform.argtypes = [_p, _sz, _sz, _sz, C.c_uint64]
        _lib.na_oracle_fill_uniform.restype = None
        _lib.na_oracle_dot_f64.argtypes = [_sz, _p, _pd, _p, _pd]
        _lib.na_oracle_dot_f64.restype = _dbl
        _lib.na_oracle_cholesky_f64.argtypes = [_sz, _p, _sz, _int, _dbl, _p]
        _lib.na_oracle_cholesky_f64.restype = _int
        _lib.na_oracle_cholesky_solve_f64.argtypes = [_sz, _p, _sz, _p, _sz, _sz]
        _lib.na_oracle_cholesky_solve_f64.restype = None
        _lib.na_oracle_lu_f64.argtypes = [_sz, _sz, _p, _sz, _p, _p]
        _lib.na_oracle_lu_f64.restype = None
        _lib.na_oracle_lu_solve_f64.argtypes = [_sz, _p, _sz, _p, _sz, _p, _sz, _sz]
        _lib.na_oracle_lu_solve_f64.restype = _int
        _lib.na_oracle_icamax_f64.argtypes = [_sz, _p, _pd]
        _lib.na_oracle_icamax_f64.restype = _sz
        for name in ("na_oracle_permute_rows_f64", "na_oracle_inv_permute_rows_f64"):
            getattr(_lib, name).argtypes = [_p, _sz, _p, _sz, _sz]
            getattr(_lib, name).restype = None
        _lib.na_oracle_lu_determinant_f64.argtypes = [_sz, _p, _sz, _sz]
        _lib.na_oracle_lu_determinant_f64.restype = _dbl
        _lib.na_oracle_try_invert_f64.argtypes = [_sz, _p, _sz, _p, _sz]
        _lib.na_oracle_try_invert_f64.restype = _int
        _lib.na_oracle_qr_f64.argtypes = [_sz, _sz, _p, _sz, _p]
        _lib.na_oracle_qr_f64.restype = None
        _lib.na_oracle_full_piv_lu_f64.argtypes = [_sz, _sz, _p, _sz, _p, _p, _p, _p]
        _lib.na_oracle_full_piv_lu_f64.restype = None
        _lib.na_oracle_col_piv_qr_f64.argtypes = [_sz, _sz, _p, _sz, _p, _p, _p]
        _lib.na_oracle_col_piv_qr_f64.restype = None
        _lib.na_oracle_qr_q_f64.argtypes = [_sz, _sz, _p, _sz, _p, _p, _sz]
        _lib.na_oracle_qr_q_f64.restype = None
        _lib.na_oracle_qr_r_f64.argtypes = [_sz, _sz, _p, _sz, _p, _p, _sz]
        _lib.na_oracle_qr_r_f64.restype = None
        _lib.na_oracle_qr_q_tr_mul_f64.argtypes = [_sz, _sz, _p, _sz, _p, _p, _sz, _sz]
        _lib.na_oracle_qr_q_tr_mul_f64.restype = None
        _lib.na_oracle_qr_solve_f64.argtypes = [_sz, _p, _sz, _p, _p, _sz, _sz]
        _lib.na_oracle_qr_solve_f64.restype = _int
        for name in ("na_oracle_solve_lower_f64", "na_oracle_solve_upper_f64"):
            getattr(_lib, name).argtypes = [_sz, _p, _sz, _p, _sz, _sz]
            getattr(_lib, name).restype = _int
        _lib.na_oracle_solve_lower_with_diag_f64.argtypes = [_sz, _p, _sz, _dbl, _p, _sz, _sz]
        _lib.na_oracle_solve_lower_with_diag_f64.restype = _int
        _lib.na_oracle_hessenberg_f64.argtypes = [_sz, _p, _sz, _p]
        _lib.na_oracle_hessenberg_f64.restype = None
        _lib.na_oracle_assemble_q_f64.argtypes = [_sz, _p, _sz, _p, _p, _sz]
        _lib.na_oracle_assemble_q_f64.restype = None
        _lib.na_oracle_symmetric_tridiagonal_f64.argtypes = [_sz, _p, _sz, _p]
        _lib.na_oracle_symmetric_tridiagonal_f64.restype = None
        _lib.na_oracle_bidiagonal_f64.argtypes = [_sz, _sz, _p, _sz, _p, _p]
        _lib.na_oracle_bidiagonal_f64.restype = _int
        for name in ("na_oracle_bidiagonal_u_f64", "na_oracle_bidiagonal_v_t_f64"):
            getattr(_lib, name).argtypes = [_sz, _sz, _p, _sz, _p, _p, _p, _sz]
            getattr(_lib, name).restype = None
    return _lib


def _f(a, dtype=np.float64) -> np.ndarray:
    """Column-major owned copy."""
    return np.array(a, dtype=dtype, order="F", copy=True)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _strides(a: np.ndarray):
    it = a.itemsize
    if a.ndim == 1:
        return a.strides[0] // it, 0
    return a.strides[0] // it, a.strides[1] // it


# ---------------------------------------------------------------------------------------------
# synthetic inputs: same counter-based generator as the C and CUDA sides
# ---------------------------------------------------------------------------------------------
def rand01(seed: int, idx: np.ndarray) -> np.ndarray:
    """Vectorised numpy twin of ``na_oracle_rand01`` (bit-identical)."""
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(((seed + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(nrows: int, ncols: int, seed: int) -> np.ndarray:
    """nrows x ncols U[0,1) matrix (column-major), element (i,j) = rand01(seed, i + j*nrows)."""
    idx = np.arange(nrows * ncols, dtype=np.uint64)
    return rand01(seed, idx).reshape((nrows, ncols), order="F")


def spd_wellcond(n: int, seed: int) -> np.ndarray:
    """(B + B^T)/2 + n*I with B ~ U[0,1): SURVEY.md §8(d) Cfg 3 (ii)."""
    b = uniform(n, n, seed)
    return np.asfortranarray((b + b.T) * 0.5 + n * np.eye(n))


# ---------------------------------------------------------------------------------------------
# reference-shaped entry points
# ---------------------------------------------------------------------------------------------
def gemm(alpha, a, b, beta, c, *, path="dispatch", nthreads=1):
    """``c.gemm(alpha, a, b, beta)`` in place on ``c`` (any strides)."""
    m, k = a.shape
    k2, n = b.shape
    assert k == k2 and c.shape == (m, n)
    rsa, csa = _strides(a); rsb, csb = _strides(b); rsc, csc = _strides(c)
    args = (m, k, n, float(alpha), _ptr(a), rsa, csa, _ptr(b), rsb, csb, float(beta), _ptr(c), rsc, csc)
    if path == "dispatch":
        lib().na_oracle_gemm_f64(*args)
    elif path == "fallback":
        lib().na_oracle_gemm_fallback_f64(*args)
    elif path == "mm":
        lib().na_oracle_dgemm_mm(*args, int(nthreads))
    else:
        raise ValueError(path)
    return c


def gemm_f32(alpha, a, b, beta, c):
    m, k = a.shape
    _, n = b.shape
    rsa, csa = _strides(a); rsb, csb = _strides(b); rsc, csc = _strides(c)
    lib().na_oracle_gemm_f32(m, k, n, float(alpha), _ptr(a), rsa, csa, _ptr(b), rsb, csb, float(beta), _ptr(c), rsc, csc)
    return c


def gemm_tr(alpha, a, b, beta, c):
    """``c.gemm_tr(alpha, a, b, beta)``: c = alpha*a^T*b + beta*c, a is k x m."""
    k, m = a.shape
    k2, n = b.shape
    assert k == k2 and c.shape == (m, n)
    rsa, csa = _strides(a); rsb, csb = _strides(b); rsc, csc = _strides(c)
    lib().na_oracle_gemm_tr_f64(m, k, n, float(alpha), _ptr(a), rsa, csa, _ptr(b), rsb, csb, float(beta), _ptr(c), rsc, csc)
    return c


def dot(x, y) -> float:
    return lib().na_oracle_dot_f64(len(x), _ptr(x), x.strides[0] // 8, _ptr(y), y.strides[0] // 8)


def cholesky(a, substitute=None):
    """``Cholesky::new`` / ``new_with_substitute``: returns the packed matrix (lower = L, strict
    upper untouched) or ``None``."""
    a = _f(a)
    n = a.shape[0]
    assert a.shape == (n, n)
    fail = C.c_size_t(0)
    rc = lib().na_oracle_cholesky_f64(n, _ptr(a), max(n, 1), 0 if substitute is None else 1,
                                      0.0 if substitute is None else float(substitute), C.addressof(fail))
    return None if rc else a


def cholesky_solve(chol, b):
    b = _f(b)
    n = chol.shape[0]
    b2 = b.reshape(n, -1, order="F") if b.ndim == 1 else b
    lib().na_oracle_cholesky_solve_f64(n, _ptr(chol), chol.strides[1] // 8, _ptr(b2), max(n, 1), b2.shape[1])
    return b


def lu(a):
    """``LU::new``: returns (packed lu, swaps[(i, i2)...] as an (len, 2) uint64 array)."""
    a = _f(a)
    m, n = a.shape
    mn = min(m, n)
    swaps = np.zeros(2 * max(mn, 1), dtype=np.uint64)
    ns = C.c_size_t(0)
    lib().na_oracle_lu_f64(m, n, _ptr(a), max(m, 1), _ptr(swaps), C.addressof(ns))
    return a, swaps[: 2 * ns.value].reshape(-1, 2).copy()


def lu_solve(lu_packed, swaps, b):
    """``LU::solve``: returns x or ``None`` (exact-zero diagonal of U)."""
    b = _f(b)
    n = lu_packed.shape[0]
    b2 = b.reshape(n, -1, order="F") if b.ndim == 1 else b
    sw = np.ascontiguousarray(swaps, dtype=np.uint64).reshape(-1)
    ok = lib().na_oracle_lu_solve_f64(n, _ptr(lu_packed), lu_packed.strides[1] // 8, _ptr(sw), len(sw) // 2,
                                      _ptr(b2), max(n, 1), b2.shape[1])
    return b if ok else None


def lu_determinant(lu_packed, swaps) -> float:
    n = lu_packed.shape[0]
    return lib().na_oracle_lu_determinant_f64(n, _ptr(lu_packed), lu_packed.strides[1] // 8, len(swaps))


def permute_rows(swaps, b, inverse=False):
    b = _f(b)
    sw = np.ascontiguousarray(swaps, dtype=np.uint64).reshape(-1)
    fn = lib().na_oracle_inv_permute_rows_f64 if inverse else lib().na_oracle_permute_rows_f64
    fn(_ptr(sw), len(sw) // 2, _ptr(b), b.shape[0], b.shape[1])
    return b


def try_inverse(a):
    a = _f(a)
    n = a.shape[0]
    out = np.zeros((n, n), order="F")
    ok = lib().na_oracle_try_invert_f64(n, _ptr(a), max(n, 1), _ptr(out), max(n, 1))
    return out if ok else None


def lu_unpack(lu_packed):
    """(L, U) of ``LU::unpack`` (unit lower trapezoid, upper trapezoid)."""
    m, n = lu_packed.shape
    mn = min(m, n)
    l = np.tril(lu_packed[:, :mn], -1) + np.eye(m, mn)
    u = np.triu(lu_packed[:mn, :])
    return l, u


def qr(a):
    """``QR::new``: returns (qr packed, diag)."""
    a = _f(a)
    m, n = a.shape
    diag = np.zeros(max(min(m, n), 1))
    lib().na_oracle_qr_f64(m, n, _ptr(a), max(m, 1), _ptr(diag))
    return a, diag[: min(m, n)]


def full_piv_lu(a):
    """``FullPivLU::new``: returns (packed lu, p swaps, q swaps) -- swaps as (len, 2) uint64 arrays."""
    a = _f(a)
    m, n = a.shape
    mn = min(m, n)
    ps = np.zeros(2 * max(mn, 1), dtype=np.uint64); qs = np.zeros(2 * max(mn, 1), dtype=np.uint64)
    np_, nq = C.c_size_t(0), C.c_size_t(0)
    lib().na_oracle_full_piv_lu_f64(m, n, _ptr(a), max(m, 1), _ptr(ps), C.addressof(np_), _ptr(qs), C.addressof(nq))
    return a, ps[: 2 * np_.value].reshape(-1, 2).copy(), qs[: 2 * nq.value].reshape(-1, 2).copy()


def col_piv_qr(a):
    """``ColPivQR::new``: returns (packed col_piv_qr, diag, p swaps)."""
    a = _f(a)
    m, n = a.shape
    mn = min(m, n)
    diag = np.zeros(max(mn, 1)); ps = np.zeros(2 * max(mn, 1), dtype=np.uint64)
    np_ = C.c_size_t(0)
    lib().na_oracle_col_piv_qr_f64(m, n, _ptr(a), max(m, 1), _ptr(diag), _ptr(ps), C.addressof(np_))
    return a, diag[:mn], ps[: 2 * np_.value].reshape(-1, 2).copy()


def qr_q(qr_packed, diag):
    m, n = qr_packed.shape
    mn = min(m, n)
    q = np.zeros((m, mn), order="F")
    lib().na_oracle_qr_q_f64(m, n, _ptr(qr_packed), max(m, 1), _ptr(np.ascontiguousarray(diag)), _ptr(q), max(m, 1))
    return q


def qr_r(qr_packed, diag):
    m, n = qr_packed.shape
    mn = min(m, n)
    r = np.zeros((mn, n), order="F")
    lib().na_oracle_qr_r_f64(m, n, _ptr(qr_packed), max(m, 1), _ptr(np.ascontiguousarray(diag)), _ptr(r), max(mn, 1))
    return r


def qr_q_tr_mul(qr_packed, diag, b):
    b = _f(b)
    m, n = qr_packed.shape
    b2 = b.reshape(m, -1, order="F") if b.ndim == 1 else b
    lib().na_oracle_qr_q_tr_mul_f64(m, n, _ptr(qr_packed), max(m, 1), _ptr(np.ascontiguousarray(diag)), _ptr(b2), max(m, 1), b2.shape[1])
    return b


def qr_solve(qr_packed, diag, b):
    b = _f(b)
    n = qr_packed.shape[0]
    b2 = b.reshape(n, -1, order="F") if b.ndim == 1 else b
    ok = lib().na_oracle_qr_solve_f64(n, _ptr(qr_packed), max(n, 1), _ptr(np.ascontiguousarray(diag)), _ptr(b2), max(n, 1), b2.shape[1])
    return b if ok else None


def solve_lower(a, b):
    b = _f(b); a = _f(a); n = a.shape[0]
    b2 = b.reshape(n, -1, order="F") if b.ndim == 1 else b
    ok = lib().na_oracle_solve_lower_f64(n, _ptr(a), max(n, 1), _ptr(b2), max(n, 1), b2.shape[1])
    return b if ok else None


def solve_upper(a, b):
    b = _f(b); a = _f(a); n = a.shape[0]
    b2 = b.reshape(n, -1, order="F") if b.ndim == 1 else b
    ok = lib().na_oracle_solve_upper_f64(n, _ptr(a), max(n, 1), _ptr(b2), max(n, 1), b2.shape[1])
    return b if ok else None


# ---------------------------------------------------------------------------------------------
# two-sided Householder reductions
# ---------------------------------------------------------------------------------------------
def hessenberg(a):
    """``Hessenberg::new``: returns (packed hess, subdiag)."""
    a = _f(a)
    n = a.shape[0]
    assert a.shape == (n, n) and n > 0
    sub = np.zeros(max(n - 1, 1))
    lib().na_oracle_hessenberg_f64(n, _ptr(a), n, _ptr(sub))
    return a, sub[: n - 1]


def assemble_q(m, signs):
    """``householder::assemble_q`` (Hessenberg::q, SymmetricTridiagonal::q)."""
    m = _f(m)
    n = m.shape[0]
    q = np.zeros((n, n), order="F")
    s = np.ascontiguousarray(np.append(np.asarray(signs, dtype=np.float64), 0.0))
    lib().na_oracle_assemble_q_f64(n, _ptr(m), n, _ptr(s), _ptr(q), n)
    return q


def hessenberg_h(hess, subdiag):
    """``Hessenberg::h``: upper Hessenberg part with |subdiag| on the first subdiagonal."""
    n = hess.shape[0]
    h = np.triu(hess, -1)
    if n > 1:
        h[np.arange(1, n), np.arange(n - 1)] = np.abs(subdiag)
    return np.asfortranarray(h)


def symmetric_tridiagonal(a):
    """``SymmetricTridiagonal::new``: returns (packed tri, off_diagonal); only the lower triangle is read / written."""
    a = _f(a)
    n = a.shape[0]
    assert a.shape == (n, n) and n > 0
    off = np.zeros(max(n - 1, 1))
    lib().na_oracle_symmetric_tridiagonal_f64(n, _ptr(a), n, _ptr(off))
    return a, off[: n - 1]


def bidiagonal(a):
    """``Bidiagonal::new``: returns (packed uv, diagonal, off_diagonal, upper_diagonal)."""
    a = _f(a)
    m, n = a.shape
    mn = min(m, n)
    assert mn > 0
    d = np.zeros(mn); e = np.zeros(max(mn - 1, 1))
    upper = lib().na_oracle_bidiagonal_f64(m, n, _ptr(a), m, _ptr(d), _ptr(e))
    return a, d, e[: mn - 1], bool(upper)


def bidiagonal_u(uv, d, e):
    m, n = uv.shape
    mn = min(m, n)
    uv = _f(uv)
    u = np.zeros((m, mn), order="F")
    e1 = np.ascontiguousarray(np.append(np.asarray(e, dtype=np.float64), 0.0))
    lib().na_oracle_bidiagonal_u_f64(m, n, _ptr(uv), m, _ptr(np.ascontiguousarray(d)), _ptr(e1), _ptr(u), m)
    return u


def bidiagonal_v_t(uv, d, e):
    m, n = uv.shape
    mn = min(m, n)
    uv = _f(uv)
    vt = np.zeros((mn, n), order="F")
    e1 = np.ascontiguousarray(np.append(np.asarray(e, dtype=np.float64), 0.0))
    lib().na_oracle_bidiagonal_v_t_f64(m, n, _ptr(uv), m, _ptr(np.ascontiguousarray(d)), _ptr(e1), _ptr(vt), mn)
    return vt


def bidiagonal_d(d, e, upper):
    """``Bidiagonal::d``."""
    mn = len(d)
    res = np.diag(np.abs(d))
    if mn > 1:
        idx = np.arange(mn - 1)
        if upper:
            res[idx, idx + 1] = np.abs(e)
        else:
            res[idx + 1, idx] = np.abs(e)
    return np.asfortranarray(res)
