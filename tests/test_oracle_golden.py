"""CPU: pins the oracle (oracle/nalgebra_oracle.c) against every known-answer test the reference's
own test-suite holds for the hot path (tests/golden/nalgebra_kats.json), re-expresses the
reference's property tests at its own sizes (PROPTEST_MATRIX_DIM = 1..=20) and cross-checks the
restatement against LAPACK as a second opinion."""
import numpy as np
import pytest
import scipy.linalg as sl

from helpers import EPS, bench_spd, load_kats, mat, random_sdp, relative_eq

K = load_kats()


def test_rng_twins_are_bit_identical(oracle):
    x = oracle.uniform(37, 11, 42)
    lib = oracle.lib()
    for (i, j) in [(0, 0), (36, 10), (5, 7), (17, 3)]:
        assert x[i, j] == lib.na_oracle_rand01(42, i + j * 37)
    assert 0.0 <= x.min() and x.max() < 1.0
    y = np.zeros((37, 11), order="F")
    lib.na_oracle_fill_uniform(y.ctypes.data, 37, 11, 37, 42)
    assert np.array_equal(x, y)


# ---- GEMM family ---------------------------------------------------------------------------------
def test_gemm_doc(oracle):
    k = K["gemm_doc"]
    m1, m2, m3 = mat(k["mat1"]), mat(k["mat2"]), mat(k["mat3"])
    expected = (m2 @ m3) * 10.0 + m1 * 5.0
    oracle.gemm(k["alpha"], m2, m3, k["beta"], m1)
    assert relative_eq(m1, expected, EPS)


def test_gemm_tr_doc_is_exact(oracle):
    k = K["gemm_tr_doc"]
    m1, m2, m3 = mat(k["mat1"]), mat(k["mat2"]), mat(k["mat3"])
    # expected = mat2.transpose() * mat3 * 10.0 + mat1 * 5.0, evaluated with the reference's own
    # arithmetic (static-size product = gemv fallback, then scalar ops)
    prod = np.zeros((2, 4), order="F")
    oracle.gemm(1.0, np.asfortranarray(m2.T), m3, 0.0, prod, path="fallback")
    expected = prod * 10.0 + m1 * 5.0
    oracle.gemm_tr(k["alpha"], m2, m3, k["beta"], m1)
    assert np.array_equal(m1, expected)          # assert_eq! in the doc-test


def test_level1_docs(oracle):
    k = K["dot_doc"]
    assert oracle.dot(np.array(k["x"]), np.array(k["y"])) == k["expected"]
    # axcpy / axpy through a 1-column gemm fallback: y = a*x*c + b*y
    k = K["axcpy_doc"]
    y = np.array(k["y"]).reshape(3, 1, order="F").copy(order="F")
    oracle.gemm(k["a"], np.array(k["x"]).reshape(3, 1, order="F"), np.array([[k["c"]]]), k["b"], y, path="fallback")
    assert np.array_equal(y[:, 0], np.array(k["expected"]))
    k = K["axpy_doc"]
    y = np.array(k["y"]).reshape(3, 1, order="F").copy(order="F")
    oracle.gemm(k["a"], np.array(k["x"]).reshape(3, 1, order="F"), np.array([[1.0]]), k["b"], y, path="fallback")
    assert np.array_equal(y[:, 0], np.array(k["expected"]))


def test_simple_mul(oracle):
    k = K["simple_mul"]
    out = np.full((2, 4), np.nan, order="F")
    oracle.gemm(1.0, mat(k["a"]), mat(k["b"]), 0.0, out)
    assert np.array_equal(out, mat(k["expected"]))


def test_empty_matrices(oracle):
    for (m, kk, n) in K["empty_matrix_mul_matrix"]["shapes"]:
        out = np.full((m, n), np.nan, order="F")
        oracle.gemm(1.0, np.zeros((m, kk), order="F"), np.zeros((kk, n), order="F"), 0.0, out)
        assert np.array_equal(out, np.zeros((m, n)))
    k = K["empty_matrix_gemm"]
    for (m, kk, n) in k["shapes"]:
        for path in ("dispatch", "fallback", "mm"):
            out = np.full((m, n), k["c_init"], order="F")
            oracle.gemm(k["alpha"], np.zeros((m, kk), order="F"), np.zeros((kk, n), order="F"), k["beta"], out, path=path)
            assert np.array_equal(out, np.full((m, n), k["expected_fill"]))
    k = K["empty_matrix_gemm_tr"]
    out = np.full((3, 4), k["c_init"], order="F")
    oracle.gemm_tr(k["alpha"], np.zeros(tuple(k["shape_a"]), order="F"), np.zeros(tuple(k["shape_b"]), order="F"), k["beta"], out)
    assert np.array_equal(out, np.full((3, 4), k["expected_fill"]))


def test_gemm_order_and_beta(oracle):
    k = K["gemm_noncommutative"]
    m1, m2 = mat(k["m1"]), mat(k["m2"])
    res = np.zeros((2, 2), order="F")
    oracle.gemm(1.0, m1, m2, 0.0, res)
    assert np.array_equal(res, np.eye(2))
    res = np.asfortranarray(np.eye(2))
    oracle.gemm(k["k"], m1, m2, -k["k"], res)
    assert np.array_equal(res, np.zeros((2, 2)))


def test_gemm_beta_zero_never_reads_c(oracle):
    a, b = oracle.uniform(9, 8, 1), oracle.uniform(8, 7, 2)
    for path in ("dispatch", "fallback", "mm"):
        c = np.full((9, 7), np.nan, order="F")
        oracle.gemm(2.0, a, b, 0.0, c, path=path)
        assert np.allclose(c, 2.0 * a @ b, rtol=1e-14)


@pytest.mark.parametrize("shape", [(6, 6, 6), (13, 14, 15), (64, 70, 9), (129, 257, 65), (300, 300, 300)])
def test_matrixmultiply_standin_vs_fallback_and_views(oracle, shape):
    m, k, n = shape
    a, b, c0 = oracle.uniform(m, k, 1) - 0.5, oracle.uniform(k, n, 2) - 0.5, oracle.uniform(m, n, 3)
    ref = 1.5 * a @ b + 0.5 * c0
    tol = 4 * k * EPS * np.linalg.norm(a) * np.linalg.norm(b)
    for path in ("mm", "fallback"):
        c = c0.copy(order="F")
        oracle.gemm(1.5, a, b, 0.5, c, path=path)
        assert np.abs(c - ref).max() <= tol
    # strided views: transposed A, row-major B, strided C
    c = np.zeros((2 * m, 3 * n))[::2, ::3]
    c[...] = c0
    oracle.gemm(1.5, np.ascontiguousarray(a.T).T, np.ascontiguousarray(b), 0.5, c, path="mm")
    assert np.abs(c - ref).max() <= tol
    c = c0.copy(order="F")
    oracle.gemm(1.5, a, b, 0.5, c, path="mm", nthreads=3)
    assert np.abs(c - ref).max() <= tol


# ---- Cholesky ------------------------------------------------------------------------------------
def test_cholesky_with_substitute(oracle):
    k = K["cholesky_with_substitute"]
    m = mat(k["m"])
    assert oracle.cholesky(m) is None
    assert oracle.cholesky(m, substitute=k["substitute"]) is not None


@pytest.mark.parametrize("n", list(range(1, 21)))
def test_cholesky_properties(oracle, n):
    rng = np.random.default_rng(n)
    t = K["proptest_params"]["tolerances"]
    m = random_sdp(n, rng)
    l = np.tril(oracle.cholesky(m))
    assert relative_eq(m, l @ l.T, t["cholesky_recompose"])
    b = rng.random((n, 3))
    x = oracle.cholesky_solve(oracle.cholesky(m), b)
    assert relative_eq(m @ x, b, t["cholesky_solve"])
    assert np.allclose(l, np.linalg.cholesky(m), atol=1e-12)


def test_cholesky_not_positive_definite(oracle):
    rng = np.random.default_rng(0)
    m = random_sdp(8, rng)
    m[5, 5] = -1.0
    assert oracle.cholesky(m) is None
    m = bench_spd(30, rng)
    assert oracle.cholesky(m) is not None


# ---- LU ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["lu_simple", "lu_simple_with_pivot"])
def test_lu_kats(oracle, name):
    k = K[name]
    m = mat(k["m"])
    lu, sw = oracle.lu(m)
    assert oracle.lu_determinant(lu, sw) == k["determinant"]          # assert_eq!, exact
    l, u = oracle.lu_unpack(lu)
    assert relative_eq(m, oracle.permute_rows(sw, l @ u, inverse=True), k["epsilon"])


def test_matrix5_try_inverse(oracle):
    k = K["matrix5_try_inverse"]
    inv = oracle.try_inverse(mat(k["a"]))
    exp = mat(k["expected_inverse"])
    d = np.abs(inv - exp)
    assert np.all((d <= EPS) | (d <= k["max_relative"] * np.maximum(np.abs(inv), np.abs(exp))) | (np.abs(exp) < 1e-15))


@pytest.mark.parametrize("shape", [(n, n) for n in range(1, 21)] + [(3, 5), (5, 3), (4, 4), (7, 20), (20, 7)])
def test_lu_properties_and_lapack_layout(oracle, shape):
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    t = K["proptest_params"]["tolerances"]
    m = np.asfortranarray(rng.uniform(-100, 100, size=shape))
    lu, sw = oracle.lu(m)
    l, u = oracle.lu_unpack(lu)
    assert relative_eq(m, oracle.permute_rows(sw, l @ u, inverse=True), t["lu_recompose"] * 100)
    # second opinion: same pivots and packed factors as LAPACK getrf
    lu2, piv, _ = sl.lapack.dgetrf(m)
    ref_sw = [(i, int(p)) for i, p in enumerate(piv) if int(p) != i]
    assert [tuple(int(v) for v in r) for r in sw] == ref_sw
    assert np.allclose(lu, lu2, rtol=1e-10, atol=1e-10)
    if shape[0] == shape[1]:
        b = rng.uniform(-100, 100, size=(shape[0], 2))
        x = oracle.lu_solve(lu, sw, b)
        assert x is not None and np.allclose(m @ x, b, rtol=1e-6, atol=1e-6)


def test_lu_zero_column_and_singular_solve(oracle):
    a = oracle.uniform(12, 12, 6)
    a[:, 3] = 0.0
    lu, sw = oracle.lu(a)
    assert not np.isnan(lu).any()
    assert oracle.lu_solve(lu, sw, np.ones((12, 1))) is None
    assert oracle.lib().na_oracle_icamax_f64(4, np.array([1.0, -3.0, 3.0, 2.0]).ctypes.data, 1) == 1   # lowest index wins ties


# ---- QR ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(n, n) for n in range(1, 21)] + [(5, 3), (3, 5), (4, 4), (20, 7), (7, 20)])
def test_qr_properties(oracle, shape):
    rng = np.random.default_rng(shape[0] * 17 + shape[1])
    t = K["proptest_params"]["tolerances"]
    m = np.asfortranarray(rng.uniform(-100, 100, size=shape))
    qr, d = oracle.qr(m)
    q, r = oracle.qr_q(qr, d), oracle.qr_r(qr, d)
    k = min(shape)
    assert relative_eq(m, q @ r, t["qr_recompose"] * 100)
    assert np.abs(q.T @ q - np.eye(k)).max() <= t["qr_orthogonal"]
    assert np.all(np.diag(r) >= 0)                         # R[i,i] = |diag[i]|
    for i in range(k):                                     # axes are unit vectors
        assert abs(np.linalg.norm(qr[i:, i]) - 1.0) <= 1e-12 or d[i] == 0.0
    # second opinion: |R| equals LAPACK's |R|
    r2 = np.linalg.qr(m, mode="r")
    assert np.allclose(np.abs(r), np.abs(r2), rtol=1e-9, atol=1e-9)
    if shape[0] == shape[1]:
        b = rng.uniform(-100, 100, size=(shape[0], 2))
        x = oracle.qr_solve(qr, d, b)
        assert x is not None and np.allclose(m @ x, b, rtol=1e-6, atol=1e-6)
    b = rng.uniform(-1, 1, size=(shape[0], 3))
    assert np.allclose(oracle.qr_q_tr_mul(qr, d, b)[:k], q.T @ b, atol=1e-10)


def test_triangular_solves(oracle):
    rng = np.random.default_rng(5)
    a = rng.random((6, 6)) + 3 * np.eye(6)
    b = rng.random((6, 4))
    assert np.allclose(np.tril(a) @ oracle.solve_lower(a, b), b)
    assert np.allclose(np.triu(a) @ oracle.solve_upper(a, b), b)
    a[2, 2] = 0.0
    assert oracle.solve_lower(a, b) is None and oracle.solve_upper(a, b) is None
