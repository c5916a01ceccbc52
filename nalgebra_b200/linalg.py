"""Host-side mirror of nalgebra's interface for the hot path, over the C ABI.

The reference is Rust and this image has no Rust toolchain, so the host side above the C ABI is
mirrored here (the Rust crate a maintainer would add is written out under ``rust/`` and explained in
INTEGRATION.md; it cannot be compiled in this image).  Names, argument meaning and error behaviour
follow the reference:

* ``gemm(alpha, a, b, beta, c)``      -- ``Matrix::gemm``      src/base/blas.rs:729-746
* ``gemm_tr`` / ``mul_to`` / ``mul``  -- blas.rs:770-803, ops.rs:783-795, ops.rs:554-574
* ``Cholesky``                        -- src/linalg/cholesky.rs
* ``LU`` + ``PermutationSequence``    -- src/linalg/lu.rs, permutation_sequence.rs
* ``QR``                              -- src/linalg/qr.rs

A ``DMatrix<f64>`` is a numpy float64 array (any strides; ``VecStorage`` = Fortran order).  Shape
errors raise ``ValueError`` (the reference panics); numerical failure is a value (``None`` /
``False``) exactly like the reference.  All arithmetic happens in libnalgebra_b200.so on the GPU --
there is no CPU fallback; only O(n^2) accessors (``l()``, ``u()``, ``r()``, determinants) are host
code, as they are plain loops in the reference too.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import NA_NOT_PD, NA_OK, NA_SINGULAR, check


def _strides(a: np.ndarray):
    it = a.itemsize
    if a.ndim == 1:
        return a.strides[0] // it, 0
    return a.strides[0] // it, a.strides[1] // it


def _as_matrix(a, dtype=np.float64) -> np.ndarray:
    a = np.asarray(a, dtype=dtype)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError("expected a matrix")
    return a


def _owned(a) -> np.ndarray:
    """``clone_owned``: a column-major copy (VecStorage layout)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)          # a DVector is an n x 1 matrix (NOT 1 x n, which ndmin=2 would give)
    elif a.ndim != 2:
        raise ValueError("expected a matrix or a vector")
    return np.array(a, order="F", copy=True)


def kernel_launches() -> int:
    return int(_capi.lib().na_kernel_launches())


# ---------------------------------------------------------------------------------------------
# GEMM family
# ---------------------------------------------------------------------------------------------
def gemm(alpha: float, a, b, beta: float, c: np.ndarray) -> np.ndarray:
    """``c.gemm(alpha, &a, &b, beta)``: c <- alpha*a*b + beta*c in place; c is not read when
    beta == 0.  Views with arbitrary strides are accepted, like the reference."""
    a = _as_matrix(a); b = _as_matrix(b)
    if not isinstance(c, np.ndarray) or c.dtype != np.float64 or c.ndim != 2 or not c.flags.writeable:
        raise ValueError("c must be a writable float64 matrix")
    (m, k), (k2, n) = a.shape, b.shape
    if k != k2:
        raise ValueError("gemm: dimensions mismatch for multiplication.")      # blas_uninit.rs:244-247
    if c.shape != (m, n):
        raise ValueError("gemm: dimensions mismatch for addition.")            # blas_uninit.rs:248-252
    rsa, csa = _strides(a); rsb, csb = _strides(b); rsc, csc = _strides(c)
    check(_capi.lib().na_dgemm(m, k, n, float(alpha), a.ctypes.data, rsa, csa, b.ctypes.data, rsb, csb,
                               float(beta), c.ctypes.data, rsc, csc))
    return c


def gemm_tr(alpha: float, a, b, beta: float, c: np.ndarray) -> np.ndarray:
    """``c.gemm_tr(alpha, &a, &b, beta)``: c <- alpha*a^T*b + beta*c (blas.rs:770-803)."""
    a = _as_matrix(a)
    if a.shape[0] != _as_matrix(b).shape[0]:
        raise ValueError("gemm: dimensions mismatch for multiplication.")
    return gemm(alpha, a.T, b, beta, c)


def mul_to(a, b, out: np.ndarray) -> np.ndarray:
    """``a.mul_to(&b, &mut out)`` = ``out.gemm(1, a, b, 0)`` (ops.rs:783-795)."""
    return gemm(1.0, a, b, 0.0, out)


def mul(a, b) -> np.ndarray:
    """``&a * &b``: allocates an uninitialised result and calls gemm with beta = 0 (ops.rs:554-574)."""
    a = _as_matrix(a); b = _as_matrix(b)
    if a.shape[1] != b.shape[0]:
        raise ValueError("Matrix multiplication dimensions mismatch")
    out = np.empty((a.shape[0], b.shape[1]), dtype=np.float64, order="F")
    return gemm(1.0, a, b, 0.0, out)


def tr_mul(a, b) -> np.ndarray:
    """``a.tr_mul(&b)`` = a^T * b (ops.rs:674-779): same kernel with A's strides swapped."""
    a = _as_matrix(a); b = _as_matrix(b)
    if a.shape[0] != b.shape[0]:
        raise ValueError("Matrix multiplication dimensions mismatch")
    out = np.empty((a.shape[1], b.shape[1]), dtype=np.float64, order="F")
    return gemm(1.0, a.T, b, 0.0, out)


def tr_mul_to(a, b, out: np.ndarray) -> np.ndarray:
    """``a.tr_mul_to(&b, &mut out)`` (ops.rs:757-766): out <- a^T * b."""
    a = _as_matrix(a); b = _as_matrix(b)
    if a.shape[0] != b.shape[0]:
        raise ValueError("Matrix multiplication dimensions mismatch")            # ops.rs:724-729
    if out.shape != (a.shape[1], b.shape[1]):
        raise ValueError("Matrix multiplication output dimensions mismatch")      # ops.rs:730-735
    return gemm(1.0, a.T, b, 0.0, out)


# f64 is its own conjugate: the adjoint variants are the transposed ones (ops.rs:687-697, 770-779;
# blas.rs:827-860).
ad_mul = tr_mul
ad_mul_to = tr_mul_to
gemm_ad = gemm_tr


def gemv(alpha: float, a, x, beta: float, y: np.ndarray) -> np.ndarray:
    """``y.gemv(alpha, &a, &x, beta)``: y <- alpha*a*x + beta*y in place (blas.rs:421-440 -> gemv_uninit,
    blas_uninit.rs:127-177); y is not read when beta == 0."""
    a = _as_matrix(a)
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 1 or y.ndim != 1 or y.dtype != np.float64 or not y.flags.writeable:
        raise ValueError("gemv: x and y must be float64 vectors (y writable)")
    m, n = a.shape
    if x.shape[0] != n or y.shape[0] != m:
        raise ValueError("Gemv: dimensions mismatch.")                          # blas_uninit.rs:143-150
    rsa, csa = _strides(a)
    check(_capi.lib().na_dgemv(m, n, float(alpha), a.ctypes.data, rsa, csa, x.ctypes.data, x.strides[0] // 8 if n else 1,
                               float(beta), y.ctypes.data, y.strides[0] // 8 if m else 1))
    return y


def gemv_tr(alpha: float, a, x, beta: float, y: np.ndarray) -> np.ndarray:
    """``y.gemv_tr(alpha, &a, &x, beta)``: y <- alpha*a^T*x + beta*y (blas.rs:503-540): the same call on the transposed view."""
    return gemv(alpha, _as_matrix(a).T, x, beta, y)


gemv_ad = gemv_tr


def syrk_lower(alpha: float, a, beta: float, c: np.ndarray) -> np.ndarray:
    """c <- alpha * a * a^T + beta * c on the LOWER triangle (incl. diagonal) of the square c only; the
    strict upper triangle is neither read nor written.  The bench SPD recipe of the reference is
    ``M * M^T`` (benches/linalg/cholesky.rs:3-11); Cholesky::new reads only that triangle."""
    a = _as_matrix(a)
    if not isinstance(c, np.ndarray) or c.dtype != np.float64 or c.ndim != 2 or not c.flags.writeable:
        raise ValueError("c must be a writable float64 matrix")
    n, k = a.shape
    if c.shape != (n, n):
        raise ValueError("syrk: dimensions mismatch for addition.")
    if c.strides[0] != c.itemsize or (n > 1 and c.strides[1] < n * c.itemsize):
        raise ValueError("syrk: c must be column-major")
    rsa, csa = _strides(a)
    check(_capi.lib().na_dsyrk_lower(n, k, float(alpha), a.ctypes.data, rsa, csa, float(beta), c.ctypes.data,
                                     max(c.strides[1] // c.itemsize, 1) if n > 1 else max(n, 1)))
    return c


def gemm_f32(alpha: float, a, b, beta: float, c: np.ndarray) -> np.ndarray:
    """f32 twin (matrixmultiply::sgemm, blas_uninit.rs:276-291)."""
    a = _as_matrix(a, np.float32); b = _as_matrix(b, np.float32)
    (m, k), (k2, n) = a.shape, b.shape
    if k != k2 or c.shape != (m, n) or c.dtype != np.float32:
        raise ValueError("gemm: dimensions mismatch")
    rsa, csa = _strides(a); rsb, csb = _strides(b); rsc, csc = _strides(c)
    check(_capi.lib().na_sgemm(m, k, n, float(alpha), a.ctypes.data, rsa, csa, b.ctypes.data, rsb, csb,
                               float(beta), c.ctypes.data, rsc, csc))
    return c


# ---------------------------------------------------------------------------------------------
# PermutationSequence  (src/linalg/permutation_sequence.rs)
# ---------------------------------------------------------------------------------------------
class PermutationSequence:
    """``{len, ipiv}``: the non-trivial row interchanges (i, i2) in application order."""

    def __init__(self, ipiv: np.ndarray, capacity: int):
        self.ipiv = np.ascontiguousarray(ipiv, dtype=np.uint64).reshape(-1, 2)
        self.capacity = capacity

    @classmethod
    def identity(cls, n: int) -> "PermutationSequence":
        return cls(np.zeros((0, 2), dtype=np.uint64), n)

    def __len__(self) -> int:
        return self.ipiv.shape[0]

    def is_empty(self) -> bool:
        return len(self) == 0

    def append_permutation(self, i: int, i2: int) -> None:           # :84-93
        if i != i2:
            if len(self) >= self.capacity:
                raise ValueError("Maximum number of permutations exceeded.")
            self.ipiv = np.vstack([self.ipiv, np.array([[i, i2]], dtype=np.uint64)])

    def permute_rows(self, rhs: np.ndarray) -> None:                  # :97-104
        for i, i2 in self.ipiv:
            rhs[[int(i), int(i2)]] = rhs[[int(i2), int(i)]]

    def inv_permute_rows(self, rhs: np.ndarray) -> None:              # :108-116
        for i, i2 in self.ipiv[::-1]:
            rhs[[int(i), int(i2)]] = rhs[[int(i2), int(i)]]

    def permute_columns(self, rhs: np.ndarray) -> None:
        for i, i2 in self.ipiv:
            rhs[:, [int(i), int(i2)]] = rhs[:, [int(i2), int(i)]]

    def inv_permute_columns(self, rhs: np.ndarray) -> None:
        for i, i2 in self.ipiv[::-1]:
            rhs[:, [int(i), int(i2)]] = rhs[:, [int(i2), int(i)]]

    def determinant(self) -> float:                                   # :158-164
        return 1.0 if len(self) % 2 == 0 else -1.0


# ---------------------------------------------------------------------------------------------
# Cholesky  (src/linalg/cholesky.rs)
# ---------------------------------------------------------------------------------------------
class Cholesky:
    """``Cholesky{chol}``: L in the lower triangle (incl. diagonal) of ``chol``; the strict upper
    triangle is the caller's original data, never read nor written."""

    def __init__(self, chol: np.ndarray):
        self.chol = chol

    @classmethod
    def new(cls, matrix) -> "Cholesky | None":                        # :196-198
        return cls._new_internal(matrix, None)

    @classmethod
    def new_with_substitute(cls, matrix, substitute: float) -> "Cholesky | None":   # :217-219
        return cls._new_internal(matrix, substitute)

    @classmethod
    def _new_internal(cls, matrix, substitute):
        m = _owned(matrix)
        if m.shape[0] != m.shape[1]:
            raise ValueError("The input matrix must be square.")      # :222
        n = m.shape[0]
        fail = C.c_size_t(0)
        st = check(_capi.lib().na_cholesky_f64(n, m.ctypes.data, max(n, 1), 0 if substitute is None else 1,
                                               0.0 if substitute is None else float(substitute), C.addressof(fail)))
        if st == NA_NOT_PD:
            return None
        return cls(m)

    @classmethod
    def pack_dirty(cls, matrix) -> "Cholesky":
        return cls(_owned(matrix))

    def unpack(self) -> np.ndarray:                                   # :88-91
        return np.asfortranarray(np.tril(self.chol))

    def unpack_dirty(self) -> np.ndarray:
        return self.chol

    def l(self) -> np.ndarray:                                        # :104-106
        return np.asfortranarray(np.tril(self.chol))

    def l_dirty(self) -> np.ndarray:
        return self.chol

    def solve_mut(self, b: np.ndarray) -> None:                       # :122-129
        n = self.chol.shape[0]
        if b.shape[0] != n:
            raise ValueError("Cholesky solve matrix dimension mismatch.")
        x = _owned(b)
        assert x.shape[0] == n
        check(_capi.lib().na_cholesky_solve_f64(n, self.chol.ctypes.data, max(n, 1), x.ctypes.data, max(n, 1), x.shape[1]))
        b[...] = x.reshape(b.shape, order="F")

    def solve(self, b) -> np.ndarray:                                 # :134-143
        res = np.array(b, dtype=np.float64, order="F", copy=True)
        self.solve_mut(res)
        return res

    def inverse(self) -> np.ndarray:                                  # :147-153
        res = np.asfortranarray(np.eye(self.chol.shape[0]))
        self.solve_mut(res)
        return res

    def determinant(self) -> float:                                   # :157-164
        prod = 1.0
        for d in np.diag(self.chol):
            prod *= d
        return prod * prod

    def ln_determinant(self) -> float:                                # :172-185
        return float(sum(np.log(d * d) for d in np.diag(self.chol)))


# ---------------------------------------------------------------------------------------------
# LU  (src/linalg/lu.rs)
# ---------------------------------------------------------------------------------------------
class LU:
    """``LU{lu, p}``: packed factors (strict lower = L multipliers, upper incl. diagonal = U) and the
    row PermutationSequence."""

    def __init__(self, lu: np.ndarray, p: PermutationSequence):
        self.lu, self._p = lu, p

    @classmethod
    def new(cls, matrix) -> "LU":                                     # :93-122
        m = _owned(matrix)
        nrows, ncols = m.shape
        mn = min(nrows, ncols)
        swaps = np.zeros(2 * max(mn, 1), dtype=np.uint64)
        ns = C.c_size_t(0)
        check(_capi.lib().na_lu_f64(nrows, ncols, m.ctypes.data, max(nrows, 1), swaps.ctypes.data, C.addressof(ns)))
        return cls(m, PermutationSequence(swaps[: 2 * ns.value].copy(), mn))

    def lu_internal(self) -> np.ndarray:
        return self.lu

    def l(self) -> np.ndarray:                                        # :132-141
        m, n = self.lu.shape
        mn = min(m, n)
        return np.asfortranarray(np.tril(self.lu[:, :mn], -1) + np.eye(m, mn))

    def u(self) -> np.ndarray:                                        # :176-182
        m, n = self.lu.shape
        return np.asfortranarray(np.triu(self.lu[: min(m, n), :]))

    def p(self) -> PermutationSequence:                               # :186-188
        return self._p

    def unpack(self):                                                 # :192-210
        return self._p, self.l(), self.u()

    def solve_mut(self, b: np.ndarray) -> bool:                       # :242-260
        n = self.lu.shape[0]
        if b.shape[0] != n:
            raise ValueError("LU solve matrix dimension mismatch.")
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("LU solve: unable to solve a non-square system.")
        x = _owned(b)
        assert x.shape[0] == n
        sw = np.ascontiguousarray(self._p.ipiv.reshape(-1))
        st = check(_capi.lib().na_lu_solve_f64(n, self.lu.ctypes.data, max(n, 1), sw.ctypes.data, len(self._p),
                                               x.ctypes.data, max(n, 1), x.shape[1]))
        b[...] = x.reshape(b.shape, order="F")
        return st != NA_SINGULAR

    def solve(self, b):                                               # :221-236
        res = np.array(b, dtype=np.float64, order="F", copy=True)
        return res if self.solve_mut(res) else None

    def try_inverse(self):                                            # :266-280
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("LU inverse: unable to compute the inverse of a non-square matrix.")
        res = np.asfortranarray(np.eye(self.lu.shape[0]))
        return res if self.solve_mut(res) else None

    def try_inverse_to(self, out: np.ndarray) -> bool:               # :283-297
        """Writes the inverse into `out` (clobbered, like the reference, when it returns False)."""
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("LU inverse: unable to compute the inverse of a non-square matrix.")
        if out.shape != self.lu.shape:
            raise ValueError("LU inverse: mismatched output shape.")
        out[...] = np.eye(self.lu.shape[0])                           # out.fill_with_identity()
        return self.solve_mut(out)

    def l_unpack(self) -> np.ndarray:                                 # :156-171
        """Consumes the decomposition and returns L in its storage (the packed matrix is reused)."""
        m, n = self.lu.shape
        mn = min(m, n)
        res = np.asfortranarray(np.tril(self.lu[:, :mn], -1) + np.eye(m, mn))
        self.lu = None
        return res

    def determinant(self) -> float:                                   # :301-314
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("LU determinant: unable to compute the determinant of a non-square matrix.")
        res = 1.0
        for d in np.diag(self.lu):
            res *= d
        return res * self._p.determinant()

    def is_invertible(self) -> bool:                                  # :318-331
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("LU: unable to test the invertibility of a non-square matrix.")
        return bool(np.all(np.diag(self.lu) != 0.0))


def try_invert_to(matrix, out: np.ndarray) -> bool:
    """``lu::try_invert_to(matrix, out)`` (src/linalg/lu.rs:51-86): LU-factors `matrix` and writes its
    inverse into `out`; False when a diagonal entry of U is exactly zero."""
    m = _as_matrix(matrix)
    if m.shape[0] != m.shape[1]:
        raise ValueError("LU inversion: unable to invert a rectangular matrix.")   # lu.rs:61-64
    return LU.new(m).try_inverse_to(out)


# ---------------------------------------------------------------------------------------------
# QR  (src/linalg/qr.rs)
# ---------------------------------------------------------------------------------------------
class QR:
    """``QR{qr, diag}`` in nalgebra's storage: column i, rows i.. = unit Householder axis; strict
    upper = R off-diagonal; R[i,i] = |diag[i]|."""

    def __init__(self, qr: np.ndarray, diag: np.ndarray):
        self.qr, self.diag = qr, diag

    @classmethod
    def new(cls, matrix) -> "QR":                                     # :55-76
        m = _owned(matrix)
        nrows, ncols = m.shape
        mn = min(nrows, ncols)
        diag = np.zeros(max(mn, 1))
        check(_capi.lib().na_qr_f64(nrows, ncols, m.ctypes.data, max(nrows, 1), diag.ctypes.data))
        return cls(m, diag[:mn].copy())

    def qr_internal(self) -> np.ndarray:
        return self.qr

    def diag_internal(self) -> np.ndarray:
        return self.diag

    def r(self) -> np.ndarray:                                        # :81-89
        m, n = self.qr.shape
        mn = min(m, n)
        res = np.triu(self.qr[:mn, :])
        res[np.arange(mn), np.arange(mn)] = np.abs(self.diag)
        return np.asfortranarray(res)

    unpack_r = r

    def q(self) -> np.ndarray:                                        # :108-129
        m, n = self.qr.shape
        mn = min(m, n)
        q = np.zeros((m, max(mn, 1)), order="F")
        check(_capi.lib().na_qr_q_f64(m, n, self.qr.ctypes.data, max(m, 1), np.ascontiguousarray(self.diag).ctypes.data,
                                      q.ctypes.data, max(m, 1)))
        return q[:, :mn]

    def unpack(self):
        return self.q(), self.r()

    def q_tr_mul(self, rhs: np.ndarray) -> None:                      # :157-171
        m, n = self.qr.shape
        if rhs.shape[0] != m:
            raise ValueError("q_tr_mul: dimension mismatch")
        x = _owned(rhs)
        assert x.shape[0] == m
        check(_capi.lib().na_qr_q_tr_mul_f64(m, n, self.qr.ctypes.data, max(m, 1), np.ascontiguousarray(self.diag).ctypes.data,
                                             x.ctypes.data, max(m, 1), x.shape[1]))
        rhs[...] = x.reshape(rhs.shape, order="F")

    def solve_mut(self, b: np.ndarray) -> bool:                       # :204-221
        n = self.qr.shape[0]
        if b.shape[0] != n:
            raise ValueError("QR solve matrix dimension mismatch.")
        if self.qr.shape[0] != self.qr.shape[1]:
            raise ValueError("QR solve: unable to solve a non-square system.")
        x = _owned(b)
        assert x.shape[0] == n
        st = check(_capi.lib().na_qr_solve_f64(n, self.qr.ctypes.data, max(n, 1), np.ascontiguousarray(self.diag).ctypes.data,
                                               x.ctypes.data, max(n, 1), x.shape[1]))
        b[...] = x.reshape(b.shape, order="F")
        return st != NA_SINGULAR

    def solve(self, b):                                               # :186-200
        res = np.array(b, dtype=np.float64, order="F", copy=True)
        return res if self.solve_mut(res) else None

    def try_inverse(self):                                            # :262-277
        if self.qr.shape[0] != self.qr.shape[1]:
            raise ValueError("QR inverse: unable to compute the inverse of a non-square matrix.")
        res = np.asfortranarray(np.eye(self.qr.shape[0]))
        return res if self.solve_mut(res) else None

    def is_invertible(self) -> bool:                                  # :281-294
        if self.qr.shape[0] != self.qr.shape[1]:
            raise ValueError("QR: unable to test the invertibility of a non-square matrix.")
        return bool(np.all(self.diag != 0.0))


# ---------------------------------------------------------------------------------------------
# FullPivLU  (src/linalg/full_piv_lu.rs)
# ---------------------------------------------------------------------------------------------
class FullPivLU:
    """``FullPivLU{lu, p, q}``: P * matrix * Q = L U, packed like ``LU``; ``p`` row and ``q`` column swaps."""

    def __init__(self, lu: np.ndarray, p: PermutationSequence, q: PermutationSequence):
        self.lu, self._p, self._q = lu, p, q

    @classmethod
    def new(cls, matrix) -> "FullPivLU":                              # :56-91
        m = _owned(matrix)
        nrows, ncols = m.shape
        mn = min(nrows, ncols)
        ps = np.zeros(2 * max(mn, 1), dtype=np.uint64); qs = np.zeros(2 * max(mn, 1), dtype=np.uint64)
        np_, nq = C.c_size_t(0), C.c_size_t(0)
        check(_capi.lib().na_full_piv_lu_f64(nrows, ncols, m.ctypes.data, max(nrows, 1), ps.ctypes.data, C.addressof(np_),
                                             qs.ctypes.data, C.addressof(nq)))
        return cls(m, PermutationSequence(ps[: 2 * np_.value].copy(), mn), PermutationSequence(qs[: 2 * nq.value].copy(), mn))

    def lu_internal(self) -> np.ndarray:                              # :94
        return self.lu

    def l(self) -> np.ndarray:                                        # :101-111
        m, n = self.lu.shape
        mn = min(m, n)
        return np.asfortranarray(np.tril(self.lu[:, :mn], -1) + np.eye(m, mn))

    def u(self) -> np.ndarray:                                        # :115-122
        m, n = self.lu.shape
        return np.asfortranarray(np.triu(self.lu[: min(m, n), :]))

    def p(self) -> PermutationSequence:                               # :126
        return self._p

    def q(self) -> PermutationSequence:                               # :133
        return self._q

    def unpack(self):                                                 # :139-155
        return self._p, self.l(), self.u(), self._q

    def is_invertible(self) -> bool:                                  # :238-246
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("FullPivLU: unable to test the invertibility of a non-square matrix.")
        dim = self.lu.shape[0]
        return bool(self.lu[dim - 1, dim - 1] != 0.0)

    def solve_mut(self, b: np.ndarray) -> bool:                       # :189-215
        n = self.lu.shape[0]
        if b.shape[0] != n:
            raise ValueError("FullPivLU solve matrix dimension mismatch.")
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("FullPivLU solve: unable to solve a non-square system.")
        if not self.is_invertible():
            return False
        x = _owned(b)
        self._p.permute_rows(x)
        lib = _capi.lib()
        check(lib.na_tri_solve_f64(1, 0, 1, n, self.lu.ctypes.data, max(n, 1), x.ctypes.data, max(n, 1), x.shape[1]))
        check(lib.na_tri_solve_f64(0, 0, 0, n, self.lu.ctypes.data, max(n, 1), x.ctypes.data, max(n, 1), x.shape[1]))
        self._q.inv_permute_rows(x)
        b[...] = x.reshape(b.shape, order="F")
        return True

    def solve(self, b):                                               # :168-183
        res = np.array(b, dtype=np.float64, order="F", copy=True)
        return res if self.solve_mut(res) else None

    def try_inverse(self):                                            # :220-234
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("FullPivLU inverse: unable to compute the inverse of a non-square matrix.")
        res = np.asfortranarray(np.eye(self.lu.shape[0]))
        return res if self.solve_mut(res) else None

    def determinant(self) -> float:                                   # :250-270
        if self.lu.shape[0] != self.lu.shape[1]:
            raise ValueError("FullPivLU determinant: unable to compute the determinant of a non-square matrix.")
        dim = self.lu.shape[0]
        res = float(self.lu[dim - 1, dim - 1])
        if res == 0.0:
            return 0.0
        for i in range(dim - 1):
            res *= float(self.lu[i, i])
        return res * self._p.determinant() * self._q.determinant()


# ---------------------------------------------------------------------------------------------
# ColPivQR  (src/linalg/col_piv_qr.rs)
# ---------------------------------------------------------------------------------------------
class ColPivQR:
    """``ColPivQR{col_piv_qr, p, diag}``: matrix * P = Q R, storage as ``QR``."""

    def __init__(self, col_piv_qr: np.ndarray, p: PermutationSequence, diag: np.ndarray):
        self.col_piv_qr, self._p, self.diag = col_piv_qr, p, diag

    @classmethod
    def new(cls, matrix) -> "ColPivQR":                               # :59-93
        m = _owned(matrix)
        nrows, ncols = m.shape
        mn = min(nrows, ncols)
        diag = np.zeros(max(mn, 1))
        ps = np.zeros(2 * max(mn, 1), dtype=np.uint64)
        np_ = C.c_size_t(0)
        check(_capi.lib().na_col_piv_qr_f64(nrows, ncols, m.ctypes.data, max(nrows, 1), diag.ctypes.data, ps.ctypes.data, C.addressof(np_)))
        return cls(m, PermutationSequence(ps[: 2 * np_.value].copy(), mn), diag[:mn].copy())

    def col_piv_qr_internal(self) -> np.ndarray:                      # :176
        return self.col_piv_qr

    def r(self) -> np.ndarray:                                        # :97-108
        m, n = self.col_piv_qr.shape
        mn = min(m, n)
        res = np.triu(self.col_piv_qr[:mn, :])
        res[np.arange(mn), np.arange(mn)] = np.abs(self.diag)
        return np.asfortranarray(res)

    unpack_r = r

    def q(self) -> np.ndarray:                                        # :129-150
        return QR(self.col_piv_qr, self.diag).q()

    def p(self) -> PermutationSequence:                               # :154
        return self._p

    def unpack(self):                                                 # :159-173
        return self.q(), self.r(), self._p

    def q_tr_mul(self, rhs: np.ndarray) -> None:                      # :181-199
        QR(self.col_piv_qr, self.diag).q_tr_mul(rhs)

    def is_invertible(self) -> bool:                                  # :307-320
        if self.col_piv_qr.shape[0] != self.col_piv_qr.shape[1]:
            raise ValueError("ColPivQR: unable to test the invertibility of a non-square matrix.")
        return bool(np.all(self.diag != 0.0))

    def solve_mut(self, b: np.ndarray) -> bool:                       # :227-247
        n = self.col_piv_qr.shape[0]
        if b.shape[0] != n:
            raise ValueError("ColPivQR solve matrix dimension mismatch.")
        if self.col_piv_qr.shape[0] != self.col_piv_qr.shape[1]:
            raise ValueError("ColPivQR solve: unable to solve a non-square system.")
        x = _owned(b)
        st = check(_capi.lib().na_qr_solve_f64(n, self.col_piv_qr.ctypes.data, max(n, 1), np.ascontiguousarray(self.diag).ctypes.data,
                                               x.ctypes.data, max(n, 1), x.shape[1]))
        self._p.inv_permute_rows(x)
        b[...] = x.reshape(b.shape, order="F")
        return st != NA_SINGULAR

    def solve(self, b):                                               # :205-221
        res = np.array(b, dtype=np.float64, order="F", copy=True)
        return res if self.solve_mut(res) else None

    def try_inverse(self):                                            # :288-303
        if self.col_piv_qr.shape[0] != self.col_piv_qr.shape[1]:
            raise ValueError("ColPivQR inverse: unable to compute the inverse of a non-square matrix.")
        res = np.asfortranarray(np.eye(self.col_piv_qr.shape[0]))
        return res if self.solve_mut(res) else None

    def determinant(self) -> float:                                   # :324-337
        if self.col_piv_qr.shape[0] != self.col_piv_qr.shape[1]:
            raise ValueError("ColPivQR determinant: unable to compute the determinant of a non-square matrix.")
        res = 1.0
        for d in self.diag:
            res *= float(d)
        return res * self._p.determinant()


# ---------------------------------------------------------------------------------------------
# two-sided Householder reductions  (src/linalg/hessenberg.rs, symmetric_tridiagonal.rs, bidiagonal.rs)
# ---------------------------------------------------------------------------------------------
def _assemble_q(m: np.ndarray, signs: np.ndarray) -> np.ndarray:
    """``householder::assemble_q`` (householder.rs:132-152): the axes sit in column i, rows i + 1.. of ``m``.  The product
    of the reflectors is 1 (+) the ``QR::q`` of the storage one row down, so the blocked reflector application of the QR
    path builds it (one device call)."""
    n = m.shape[0]
    q = np.zeros((n, n), order="F")
    q[0, 0] = 1.0
    if n > 1:
        m = np.asfortranarray(m)
        it = m.itemsize
        sg = np.ascontiguousarray(signs, dtype=np.float64)
        check(_capi.lib().na_qr_q_f64(n - 1, n - 1, m.ctypes.data + it, n, sg.ctypes.data, q.ctypes.data + (1 + n) * it, n))
    return q


class Hessenberg:
    """``Hessenberg{hess, subdiag}``: matrix = Q H Q^T; ``hess`` holds H above the first subdiagonal and the unit
    Householder axes below it, ``subdiag`` the signed norms."""

    def __init__(self, hess: np.ndarray, subdiag: np.ndarray):
        self.hess, self.subdiag = hess, subdiag

    @classmethod
    def new(cls, hess) -> "Hessenberg":                               # :48-100
        m = _owned(hess)
        if m.shape[0] != m.shape[1]:
            raise ValueError("Cannot compute the hessenberg decomposition of a non-square matrix.")
        n = m.shape[0]
        if n == 0:
            raise ValueError("Cannot compute the hessenberg decomposition of an empty matrix.")
        sub = np.zeros(max(n - 1, 1))
        check(_capi.lib().na_hessenberg_f64(n, m.ctypes.data, n, sub.ctypes.data))
        return cls(m, sub[: n - 1].copy())

    @classmethod
    def new_with_workspace(cls, hess, work) -> "Hessenberg":           # :61-100: `work` (n entries) is the reference's scratch vector
        if np.asarray(work).shape[0] != np.asarray(hess).shape[0]:
            raise ValueError("Hessenberg: invalid workspace size.")
        return cls.new(hess)

    def hess_internal(self) -> np.ndarray:                            # :152
        return self.hess

    def h(self) -> np.ndarray:                                        # :128-140
        n = self.hess.shape[0]
        res = np.triu(self.hess, -1)
        if n > 1:
            res[np.arange(1, n), np.arange(n - 1)] = np.abs(self.subdiag)
        return np.asfortranarray(res)

    unpack_h = h                                                       # :111-123

    def q(self) -> np.ndarray:                                        # :144-146
        return _assemble_q(self.hess, self.subdiag)

    def unpack(self):                                                 # :104-108
        return self.q(), self.h()


class SymmetricTridiagonal:
    """``SymmetricTridiagonal{tri, off_diagonal}``: matrix = Q T Q^T for a symmetric matrix (only its lower triangle is
    read); ``tri`` keeps T's diagonal and the Householder axes below the first subdiagonal."""

    def __init__(self, tri: np.ndarray, off_diagonal: np.ndarray):
        self.tri, self._off = tri, off_diagonal

    @classmethod
    def new(cls, m) -> "SymmetricTridiagonal":                        # :54-95
        a = _owned(m)
        if a.shape[0] != a.shape[1]:
            raise ValueError("Unable to compute the symmetric tridiagonal decomposition of a non-square matrix.")
        n = a.shape[0]
        if n == 0:
            raise ValueError("Unable to compute the symmetric tridiagonal decomposition of an empty matrix.")
        off = np.zeros(max(n - 1, 1))
        check(_capi.lib().na_symmetric_tridiagonal_f64(n, a.ctypes.data, n, off.ctypes.data))
        return cls(a, off[: n - 1].copy())

    def internal_tri(self) -> np.ndarray:                             # :99
        return self.tri

    def diagonal(self) -> np.ndarray:                                 # :135-140
        return np.diagonal(self.tri).copy()

    def off_diagonal(self) -> np.ndarray:                             # :144-149
        return np.abs(self._off)

    def q(self) -> np.ndarray:                                        # :153-155
        return _assemble_q(self.tri, self._off)

    def unpack(self):                                                 # :105-119
        return self.q(), self.diagonal(), self.off_diagonal()

    def unpack_tridiagonal(self):                                     # :122-131
        return self.diagonal(), self.off_diagonal()

    def recompose(self) -> np.ndarray:                                # :158-171
        q = self.q()
        n = self.tri.shape[0]
        t = np.zeros((n, n), order="F")
        t[np.arange(n), np.arange(n)] = np.diagonal(self.tri)
        if n > 1:
            idx = np.arange(n - 1)
            t[idx + 1, idx] = np.abs(self._off); t[idx, idx + 1] = np.abs(self._off)
        return mul(mul(q, t), np.asfortranarray(q.T))


class Bidiagonal:
    """``Bidiagonal{uv, diagonal, off_diagonal, upper_diagonal}``: matrix = U D V^T with D bidiagonal (upper when
    nrows >= ncols, lower otherwise)."""

    def __init__(self, uv: np.ndarray, diagonal: np.ndarray, off_diagonal: np.ndarray, upper_diagonal: bool):
        self.uv, self._diag, self._off, self.upper_diagonal = uv, diagonal, off_diagonal, upper_diagonal

    @classmethod
    def new(cls, matrix) -> "Bidiagonal":                             # :74-150
        m = _owned(matrix)
        nrows, ncols = m.shape
        mn = min(nrows, ncols)
        if mn == 0:
            raise ValueError("Cannot compute the bidiagonalization of an empty matrix.")
        d = np.zeros(mn); e = np.zeros(max(mn - 1, 1))
        check(_capi.lib().na_bidiagonal_f64(nrows, ncols, m.ctypes.data, nrows, d.ctypes.data, e.ctypes.data))
        return cls(m, d, e[: mn - 1].copy(), nrows >= ncols)

    def is_upper_diagonal(self) -> bool:                              # :154-156
        return self.upper_diagonal

    def uv_internal(self) -> np.ndarray:                              # :306
        return self.uv

    def diagonal(self) -> np.ndarray:                                 # :287-292
        return np.abs(self._diag)

    def off_diagonal(self) -> np.ndarray:                             # :296-301
        return np.abs(self._off)

    def d(self) -> np.ndarray:                                        # :185-200
        mn = len(self._diag)
        res = np.zeros((mn, mn), order="F")
        res[np.arange(mn), np.arange(mn)] = np.abs(self._diag)
        if mn > 1:
            idx = np.arange(mn - 1)
            if self.upper_diagonal:
                res[idx, idx + 1] = np.abs(self._off)
            else:
                res[idx + 1, idx] = np.abs(self._off)
        return res

    @staticmethod
    def _q_of(storage: np.ndarray, signs: np.ndarray, shift: int) -> np.ndarray:
        """Product of the reflectors whose axes sit in column i, rows i + shift.. of ``storage`` (rows x cols): the first
        min(rows, cols) columns of it, as a rows x min(rows, cols) matrix."""
        storage = np.asfortranarray(storage)
        rows, cols = storage.shape
        k = min(rows, cols)
        q = np.zeros((rows, k), order="F")
        it = storage.itemsize
        sg = np.ascontiguousarray(signs, dtype=np.float64)
        if shift == 0:
            check(_capi.lib().na_qr_q_f64(rows, cols, storage.ctypes.data, rows, sg.ctypes.data, q.ctypes.data, rows))
        else:
            q[0, 0] = 1.0
            if rows > 1 and k > 1:
                check(_capi.lib().na_qr_q_f64(rows - 1, k - 1, storage.ctypes.data + it, rows, sg.ctypes.data,
                                              q.ctypes.data + (1 + rows) * it, rows))
        return q

    def u(self) -> np.ndarray:                                        # :205-240
        if self.upper_diagonal:
            return self._q_of(self.uv, self._diag, 0)
        return self._q_of(self.uv, self._off, 1)

    def v_t(self) -> np.ndarray:                                      # :244-283
        # the row axes are the column axes of the transposed storage; V = (v_t)^T is their reflector product
        st = np.asfortranarray(self.uv.T)
        if self.upper_diagonal:
            v = self._q_of(st, self._off, 1)
        else:
            v = self._q_of(st, self._diag, 0)
        return np.asfortranarray(v.T)

    def unpack(self):                                                 # :163-181
        return self.u(), self.d(), self.v_t()


# ---------------------------------------------------------------------------------------------
# triangular solves  (src/linalg/solve.rs)
# ---------------------------------------------------------------------------------------------
def _tri_solve(t, b, lower: bool, trans: bool, unit: bool):
    t = _owned(t)
    n = t.shape[0]
    x = _owned(b)
    if x.shape[0] != n:
        raise ValueError("triangular solve: dimension mismatch")
    st = check(_capi.lib().na_tri_solve_f64(int(lower), int(trans), int(unit), n, t.ctypes.data, max(n, 1),
                                            x.ctypes.data, max(n, 1), x.shape[1]))
    if st == NA_SINGULAR:
        return None
    return x.reshape(np.shape(b), order="F")


def solve_lower_triangular(t, b):
    return _tri_solve(t, b, True, False, False)


def solve_upper_triangular(t, b):
    return _tri_solve(t, b, False, False, False)


def tr_solve_lower_triangular(t, b):
    return _tri_solve(t, b, True, True, False)


def tr_solve_upper_triangular(t, b):
    return _tri_solve(t, b, False, True, False)


def solve_lower_triangular_with_diag(t, b, diag: float):
    """solve.rs:106-133: the diagonal of `t` is never read and taken to be `diag`; ``None`` when
    diag == 0.  The reference's loop is ``coeff = b[i] / diag; b[i+1..] -= coeff * t[i+1.., i]`` and it
    never stores ``coeff`` back into ``b[i]``, so what it returns is x' with
    (I + strict_lower(t) / diag) x' = b  (= diag * the solution of (strict_lower(t) + diag I) y = b).
    That is a unit-lower solve with the strict lower triangle scaled by 1/diag; the scaling is an
    O(n^2) host pass, the substitution runs on the GPU (LU's use, diag = 1, needs no scaling)."""
    if diag == 0.0:
        return None
    if diag == 1.0:
        return _tri_solve(t, b, True, False, True)
    tm = _as_matrix(t)
    return _tri_solve(np.tril(tm, -1) / float(diag), b, True, False, True)


# ---------------------------------------------------------------------------------------------
# Matrix-level entry points  (src/linalg/decomposition.rs: `m.lu()`, `m.qr()`, ... consume the matrix and
# return the factor object; determinant / try_inverse: src/linalg/determinant.rs:57, inverse.rs:137)
# ---------------------------------------------------------------------------------------------
def cholesky(m):                        # decomposition.rs:257
    return Cholesky.new(m)


def lu(m):                              # :48
    return LU.new(m)


def qr(m):                              # :57
    return QR.new(m)


def full_piv_lu(m):                     # :39
    return FullPivLU.new(m)


def col_piv_qr(m):                      # :66
    return ColPivQR.new(m)


def hessenberg(m):                      # :289
    return Hessenberg.new(m)


def symmetric_tridiagonalize(m):        # :370
    return SymmetricTridiagonal.new(m)


def bidiagonalize(m):                   # :23
    return Bidiagonal.new(m)


def determinant(m) -> float:
    """``Matrix::determinant`` (determinant.rs:20-58): the reference's closed forms up to dimension 3 (scalar host arithmetic,
    same operation order), ``LU::new(..).determinant()`` beyond."""
    a = _as_matrix(m)
    if a.shape[0] != a.shape[1]:
        raise ValueError("Unable to compute the determinant of a non-square matrix.")
    dim = a.shape[0]
    if dim == 0:
        return 1.0
    if dim == 1:
        return float(a[0, 0])
    if dim == 2:
        return float(a[0, 0] * a[1, 1] - a[1, 0] * a[0, 1])
    if dim == 3:
        (m11, m12, m13), (m21, m22, m23), (m31, m32, m33) = (tuple(float(v) for v in row) for row in a)
        minor_m12_m23 = m22 * m33 - m32 * m23
        minor_m11_m23 = m21 * m33 - m31 * m23
        minor_m11_m22 = m21 * m32 - m31 * m22
        return m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22
    return LU.new(a).determinant()


def try_inverse(m):
    """``Matrix::try_inverse`` (inverse.rs:16-150): closed forms up to dimension 4 on the host (dimensions 1-3 in the
    reference's operation order; 4 by cofactors, the reference unrolls the same expansion), ``lu::try_invert_to`` beyond.
    ``None`` when the determinant (or a pivot of U) is exactly zero."""
    a = _as_matrix(m)
    if a.shape[0] != a.shape[1]:
        raise ValueError("Unable to invert a non-square matrix.")
    dim = a.shape[0]
    if dim == 0:
        return np.zeros((0, 0), order="F")
    if dim == 1:
        return None if a[0, 0] == 0.0 else np.asfortranarray([[1.0 / float(a[0, 0])]])
    if dim == 2:
        m11, m12, m21, m22 = float(a[0, 0]), float(a[0, 1]), float(a[1, 0]), float(a[1, 1])
        det = m11 * m22 - m21 * m12
        if det == 0.0:
            return None
        return np.asfortranarray([[m22 / det, -m12 / det], [-m21 / det, m11 / det]])
    if dim == 3:
        (m11, m12, m13), (m21, m22, m23), (m31, m32, m33) = (tuple(float(v) for v in row) for row in a)
        minor_m12_m23 = m22 * m33 - m32 * m23
        minor_m11_m23 = m21 * m33 - m31 * m23
        minor_m11_m22 = m21 * m32 - m31 * m22
        det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22
        if det == 0.0:
            return None
        return np.asfortranarray([
            [minor_m12_m23 / det, (m13 * m32 - m33 * m12) / det, (m12 * m23 - m22 * m13) / det],
            [-minor_m11_m23 / det, (m11 * m33 - m31 * m13) / det, (m13 * m21 - m23 * m11) / det],
            [minor_m11_m22 / det, (m12 * m31 - m32 * m11) / det, (m11 * m22 - m21 * m12) / det]])
    if dim == 4:
        cof = np.empty((4, 4))
        for i in range(4):
            for j in range(4):
                minor = np.delete(np.delete(a, i, axis=0), j, axis=1)
                cof[i, j] = (-1.0) ** (i + j) * determinant(minor)
        det = float(np.dot(a[0, :], cof[0, :]))
        if det == 0.0:
            return None
        return np.asfortranarray(cof.T / det)
    return LU.new(a).try_inverse()
