"""FullPivLU and ColPivQR: the oracle pinned on the reference's own tests
(/root/reference/tests/linalg/full_piv_lu.rs:3-42, tests/linalg/col_piv_qr.rs:5-23 and their property tests at
PROPTEST_MATRIX_DIM sizes) on the CPU; the CUDA path against the oracle through the C ABI under -m gpu.
FullPivLU reproduces the reference's arithmetic exactly, so its packed factors and both permutation sequences are
compared bit for bit; ColPivQR sums its dot products in another order and is compared to 1e-10 with identical pivots."""
import numpy as np
import pytest

from helpers import EPS, relative_eq

M_SIMPLE = np.asfortranarray(np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]]))
M_PIVOT = np.asfortranarray(np.array([[0.0, -1, 2], [-1, 2, -1], [2, -1, 0]]))
M_COLPIV = np.asfortranarray(np.array([[1.0, -1, 2, 1], [-1, 3, -1, -1], [3, -5, 5, 3], [1, 2, 1, -2]]))
SHAPES = [(n, n) for n in range(1, 21)] + [(3, 5), (5, 3), (4, 3), (7, 20), (20, 7), (33, 33), (64, 64), (100, 64), (64, 100), (130, 130)]


def _apply(swaps, x, cols=False, inverse=False):
    for i, i2 in (swaps[::-1] if inverse else swaps):
        i, i2 = int(i), int(i2)
        if cols: x[:, [i, i2]] = x[:, [i2, i]]
        else: x[[i, i2]] = x[[i2, i]]


def _unpack_lu(lu):
    m, n = lu.shape
    mn = min(m, n)
    return np.tril(lu[:, :mn], -1) + np.eye(m, mn), np.triu(lu[:mn, :])


# ---- CPU: the oracle against the reference's tests ---------------------------------------------------------------
@pytest.mark.parametrize("mat,det", [(M_SIMPLE, 4.0), (M_PIVOT, -4.0)])
def test_oracle_full_piv_lu_kats(oracle, mat, det):
    lu, p, q = oracle.full_piv_lu(mat)
    d = lu[2, 2] * lu[0, 0] * lu[1, 1] * (-1.0) ** (len(p) + len(q))          # full_piv_lu.rs:250-270
    assert d == det                                                          # assert_eq! in the reference
    l, u = _unpack_lu(lu)
    rec = l @ u
    _apply(p, rec, inverse=True); _apply(q, rec, cols=True, inverse=True)
    assert relative_eq(mat, rec, 1e-7)


def test_oracle_col_piv_qr_kat(oracle):
    qr, diag, p = oracle.col_piv_qr(M_COLPIV)
    assert abs(np.prod(diag) * (-1.0) ** len(p)) <= 1e-7                      # determinant ~ 0 (col_piv_qr.rs:324-337)
    q = oracle.qr_q(qr, diag); r = np.triu(qr); r[np.arange(4), np.arange(4)] = np.abs(diag)
    rec = q @ r
    _apply(p, rec, cols=True, inverse=True)
    assert relative_eq(M_COLPIV, rec, 1e-7)


@pytest.mark.parametrize("shape", SHAPES[:26])
def test_oracle_pivoted_properties(oracle, shape):
    m, n = shape
    a = oracle.uniform(m, n, 11) * 200.0 - 100.0
    lu, p, q = oracle.full_piv_lu(a)
    l, u = _unpack_lu(lu)
    rec = l @ u
    _apply(p, rec, inverse=True); _apply(q, rec, cols=True, inverse=True)
    assert relative_eq(a, rec, 1e-7)                                         # tests/linalg/full_piv_lu.rs proptest
    assert np.abs(np.tril(lu[:, : min(m, n)], -1)).max(initial=0.0) <= 1.0    # complete pivoting: |l_ij| <= 1
    qr, diag, pp = oracle.col_piv_qr(a)
    mn = min(m, n)
    qm = oracle.qr_q(qr, diag); r = np.triu(qr[:mn, :]); r[np.arange(mn), np.arange(mn)] = np.abs(diag)
    rec = qm @ r
    _apply(pp, rec, cols=True, inverse=True)
    assert relative_eq(a, rec, 1e-7)
    assert np.abs(qm.T @ qm - np.eye(mn)).max() <= 1e-7                       # q.is_orthogonal(1.0e-7)


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("mat,det", [(M_SIMPLE, 4.0), (M_PIVOT, -4.0)])
def test_full_piv_lu_kats(nab, mat, det):
    lu = nab.FullPivLU.new(mat)
    assert lu.determinant() == det
    p, l, u, q = lu.unpack()
    rec = l @ u
    p.inv_permute_rows(rec); q.inv_permute_columns(rec)
    assert relative_eq(mat, rec, 1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + [(257, 300), (700, 512), (1000, 1000)])
def test_full_piv_lu_bit_exact_vs_oracle(nab, oracle, shape):
    m, n = shape
    a = oracle.uniform(m, n, 12) - 0.4
    got = nab.FullPivLU.new(a)
    lu_ref, p_ref, q_ref = oracle.full_piv_lu(a)
    assert np.array_equal(got.p().ipiv, p_ref) and np.array_equal(got.q().ipiv, q_ref)
    assert np.array_equal(got.lu_internal(), lu_ref)                          # same arithmetic, same bits


@pytest.mark.gpu
def test_full_piv_lu_zero_tail_ties_and_solve(nab, oracle):
    a = oracle.uniform(40, 40, 5) - 0.5
    a[:, 7] = 0.0; a[:, 21] = a[:, 3]                                          # rank 38: the elimination stops early (:73-76)
    got = nab.FullPivLU.new(a); lu_ref, p_ref, q_ref = oracle.full_piv_lu(a)
    assert np.array_equal(got.lu_internal(), lu_ref) and np.array_equal(got.p().ipiv, p_ref) and np.array_equal(got.q().ipiv, q_ref)
    assert not got.is_invertible() and got.solve(np.ones((40, 1))) is None and got.determinant() == 0.0
    t = np.asfortranarray(np.array([[1.0, -2, 2], [2, 2, -2], [-2, 1, 2]]))    # |x| ties: the first in column-major order wins
    got = nab.FullPivLU.new(t); lu_ref, p_ref, q_ref = oracle.full_piv_lu(t)
    assert np.array_equal(got.lu_internal(), lu_ref) and np.array_equal(got.p().ipiv, p_ref) and np.array_equal(got.q().ipiv, q_ref)
    a = oracle.uniform(300, 300, 6) - 0.5; b = oracle.uniform(300, 4, 7)
    lu = nab.FullPivLU.new(a)
    x = lu.solve(b)
    assert x is not None and np.abs(a @ x - b).max() <= 1e-9
    inv = lu.try_inverse()
    assert np.abs(a @ inv - np.eye(300)).max() <= 1e-8
    assert abs(lu.determinant() / np.linalg.det(a) - 1.0) <= 1e-8
    assert nab.FullPivLU.new(np.zeros((0, 0), order="F")).lu_internal().shape == (0, 0)


@pytest.mark.gpu
def test_col_piv_qr_kat(nab):
    c = nab.ColPivQR.new(M_COLPIV)
    assert abs(c.determinant()) <= 1e-7
    q, r, p = c.unpack()
    rec = q @ r
    p.inv_permute_columns(rec)
    assert relative_eq(M_COLPIV, rec, 1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + [(257, 300), (700, 512), (1000, 600), (2600, 200), (4200, 70), (8192, 40), (9000, 33)])   # > 2048 rows: CTA-per-column sweep
def test_col_piv_qr_vs_oracle(nab, oracle, shape):
    m, n = shape
    a = oracle.uniform(m, n, 13) - 0.4
    got = nab.ColPivQR.new(a)
    qr_ref, diag_ref, p_ref = oracle.col_piv_qr(a)
    assert np.array_equal(got.p().ipiv, p_ref)
    assert np.abs(got.col_piv_qr_internal() - qr_ref).max(initial=0.0) <= 1e-10
    assert np.abs(got.diag - diag_ref).max(initial=0.0) <= 1e-10
    mn = min(m, n)
    q, r, p = got.unpack()
    rec = q @ r
    p.inv_permute_columns(rec)
    assert np.linalg.norm(rec - a) <= 10 * max(m, n) * EPS * np.linalg.norm(a)
    assert np.linalg.norm(q.T @ q - np.eye(mn)) <= 10 * max(m, n) * EPS


@pytest.mark.gpu
def test_col_piv_qr_solve_inverse_and_zero_column(nab, oracle):
    a = oracle.uniform(200, 200, 8) - 0.5; b = oracle.uniform(200, 3, 9)
    c = nab.ColPivQR.new(a)
    x = c.solve(b)
    assert x is not None and np.abs(a @ x - b).max() <= 1e-9
    assert np.abs(a @ c.try_inverse() - np.eye(200)).max() <= 1e-8
    assert abs(c.determinant() / np.linalg.det(a) - 1.0) <= 1e-8
    a = oracle.uniform(30, 12, 8) - 0.5; a[:, 4] = 0.0
    got = nab.ColPivQR.new(a); qr_ref, diag_ref, p_ref = oracle.col_piv_qr(a)
    assert np.array_equal(got.p().ipiv, p_ref) and np.abs(got.col_piv_qr_internal() - qr_ref).max() <= 1e-12
    assert got.diag[-1] == 0.0 == diag_ref[-1]                                 # the zero column ends up last and is not reflected
