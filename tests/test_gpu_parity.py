"""GPU (-m gpu): parity of the CUDA path with the oracle, through the C ABI.

* seeded inputs shared bit-for-bit with the oracle (counter-based generator), sizes the oracle
  finishes in seconds;
* the reference's own KATs (tests/golden/nalgebra_kats.json) and its property tests at its sizes;
* BASELINE.json's full sizes through size-independent properties (residuals, orthogonality,
  linearity), evaluated on the device.

Tolerances are the north_star's: GEMM max error <= 4*k*eps*|A||B|; factorization residuals and
orthogonality <= 10*n*eps; LU pivots bit-exact with the oracle."""
import ctypes as C

import numpy as np
import pytest

from helpers import EPS, bench_spd, gemm_tol, load_kats, mat, random_sdp, relative_eq

pytestmark = pytest.mark.gpu
K = load_kats()


@pytest.fixture(scope="module")
def L(nab):
    from nalgebra_b200 import _capi
    lib = _capi.lib()
    _capi.check(lib.na_init(0))
    return lib


def test_native_library_is_the_path(nab, L):
    n0 = nab.kernel_launches()
    a = np.ones((64, 64), order="F")
    nab.gemm(1.0, a, a, 0.0, np.empty((64, 64), order="F"))
    assert nab.kernel_launches() > n0          # our kernels ran; there is no other path


# ---- GEMM ------------------------------------------------------------------------------------------
def test_gemm_kats(nab):
    k = K["gemm_doc"]
    m1, m2, m3 = mat(k["mat1"]), mat(k["mat2"]), mat(k["mat3"])
    expected = (m2 @ m3) * 10.0 + m1 * 5.0
    nab.gemm(k["alpha"], m2, m3, k["beta"], m1)
    assert np.allclose(m1, expected, rtol=1e-14, atol=0)
    k = K["gemm_tr_doc"]
    m1, m2, m3 = mat(k["mat1"]), mat(k["mat2"]), mat(k["mat3"])
    expected = (m2.T @ m3) * 10.0 + m1 * 5.0
    nab.gemm_tr(k["alpha"], m2, m3, k["beta"], m1)
    assert np.allclose(m1, expected, rtol=1e-14, atol=0)
    k = K["simple_mul"]
    assert np.array_equal(nab.mul(mat(k["a"]), mat(k["b"])), mat(k["expected"]))     # integer-valued: exact
    k = K["gemm_noncommutative"]
    res = np.zeros((2, 2), order="F")
    nab.gemm(1.0, mat(k["m1"]), mat(k["m2"]), 0.0, res)
    assert np.array_equal(res, np.eye(2))
    res = np.asfortranarray(np.eye(2))
    nab.gemm(k["k"], mat(k["m1"]), mat(k["m2"]), -k["k"], res)
    assert np.array_equal(res, np.zeros((2, 2)))


def test_gemm_empty_and_k_zero(nab):
    for (m, kk, n) in K["empty_matrix_mul_matrix"]["shapes"]:
        assert np.array_equal(nab.mul(np.zeros((m, kk)), np.zeros((kk, n))), np.zeros((m, n)))
    k = K["empty_matrix_gemm"]
    for (m, kk, n) in k["shapes"]:
        out = np.full((m, n), k["c_init"], order="F")
        nab.gemm(k["alpha"], np.zeros((m, kk), order="F"), np.zeros((kk, n), order="F"), k["beta"], out)
        assert np.array_equal(out, np.full((m, n), k["expected_fill"]))
        out32 = np.full((m, n), k["c_init"], order="F", dtype=np.float32)
        nab.gemm_f32(k["alpha"], np.zeros((m, kk), dtype=np.float32), np.zeros((kk, n), dtype=np.float32), k["beta"], out32)
        assert np.array_equal(out32, np.full((m, n), k["expected_fill"], dtype=np.float32))
    k = K["empty_matrix_gemm_tr"]
    out = np.full((3, 4), k["c_init"], order="F")
    nab.gemm_tr(k["alpha"], np.zeros(tuple(k["shape_a"]), order="F"), np.zeros(tuple(k["shape_b"]), order="F"), k["beta"], out)
    assert np.array_equal(out, np.full((3, 4), k["expected_fill"]))
    assert nab.mul(np.zeros((0, 5)), np.zeros((5, 3))).shape == (0, 3)


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 4), (6, 6, 6), (7, 5, 3), (13, 14, 15), (17, 33, 9), (64, 64, 64),
                                   (127, 129, 130), (128, 16, 128), (200, 1000, 50), (513, 257, 255)])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (1.5, 0.5)])
def test_gemm_vs_oracle_all_layouts(nab, oracle, shape, alpha, beta):
    m, k, n = shape
    a, b, c0 = oracle.uniform(m, k, 1) - 0.5, oracle.uniform(k, n, 2) - 0.5, oracle.uniform(m, n, 3)
    ref = c0.copy(order="F")
    oracle.gemm(alpha, a, b, beta, ref)
    tol = gemm_tol(a, b, k)
    for ta in (False, True):
        for tb in (False, True):
            for tc in (False, True):
                A = a.copy(order="C" if ta else "F"); B = b.copy(order="C" if tb else "F")
                Cm = c0.copy(order="C" if tc else "F")
                if beta == 0.0:
                    Cm[...] = np.nan                       # C may be uninitialised memory when beta == 0
                nab.gemm(alpha, A, B, beta, Cm)
                assert np.abs(Cm - ref).max() <= tol, (ta, tb, tc)


def test_gemm_config0_1024_cubed(nab, oracle):
    """BASELINE configs[0]: DMatrix<f64> 1024x1024 gemm, (alpha, beta) in {(1,0), (1.5,0.5)}, plus a transposed view."""
    n = 1024
    a, b, c0 = oracle.uniform(n, n, 1), oracle.uniform(n, n, 2), oracle.uniform(n, n, 3)
    tol = gemm_tol(a, b, n)
    for alpha, beta in [(1.0, 0.0), (1.5, 0.5)]:
        ref = c0.copy(order="F"); oracle.gemm(alpha, a, b, beta, ref)
        got = c0.copy(order="F"); nab.gemm(alpha, a, b, beta, got)
        assert np.abs(got - ref).max() <= tol
    ref = np.zeros((n, n), order="F"); oracle.gemm_tr(1.0, a, b, 0.0, ref)
    got = np.zeros((n, n), order="F"); nab.gemm_tr(1.0, a, b, 0.0, got)
    assert np.abs(got - ref).max() <= tol
    assert np.abs(nab.tr_mul(a, b) - ref).max() <= tol


def test_gemm_strided_views_and_odd_offsets(nab, oracle):
    big = oracle.uniform(300, 300, 5)
    a = big[3:150:2, 1:200:3]
    b = big[5:5 + a.shape[1], 7:90]
    ref = a @ b
    cv = np.zeros((2 * a.shape[0], 3 * b.shape[1]))[::2, ::3]
    nab.gemm(1.0, a, b, 0.0, cv)
    assert np.abs(cv - ref).max() <= gemm_tol(a, b, a.shape[1])
    a = big[1:100, 1:80]; b = big[1:80, 3:50]               # 8-byte (not 16-byte) aligned views
    out = np.zeros((99, 47), order="F")
    nab.gemm(1.0, a, b, 0.0, out)
    assert np.abs(out - a @ b).max() <= gemm_tol(a, b, 79)
    a = big[::-1, :][:50, :40]                               # negative row stride
    out = np.zeros((50, 30), order="F")
    nab.gemm(1.0, a, big[:40, :30], 0.0, out)
    assert np.abs(out - a @ big[:40, :30]).max() <= gemm_tol(a, big[:40, :30], 40)


def test_gemm_host_pipelined_path(nab):
    """Host-pointer na_dgemm above the pipelining threshold (m, n >= 2048, k >= 512): non-uniform chunk grid, shell order,
    K-slabs of the first chunk (k >= 2048), C uploaded only when beta != 0 (NaN-filled otherwise)."""
    rng = np.random.default_rng(7)
    for (m, k, n) in [(2200, 2050, 2100), (2048, 600, 2304), (8320, 1100, 2176)]:
        a = np.asfortranarray(rng.random((m, k)) - 0.5); b = np.asfortranarray(rng.random((k, n)) - 0.5)
        c0 = np.asfortranarray(rng.random((m, n)))
        for (alpha, beta) in [(1.0, 0.0), (1.5, 0.5)]:
            c = c0.copy(order="F")
            if beta == 0.0:
                c[...] = np.nan
            nab.gemm(alpha, a, b, beta, c)
            ref = alpha * (a @ b) + (beta * c0 if beta != 0.0 else 0.0)
            assert np.abs(c - ref).max() <= gemm_tol(a, b, k), (m, k, n, alpha, beta)


def test_sgemm_ffma_small_path_vs_oracle(nab, oracle):
    a = (oracle.uniform(130, 70, 1) - 0.5).astype(np.float32)
    b = (oracle.uniform(70, 90, 2) - 0.5).astype(np.float32)
    c0 = oracle.uniform(130, 90, 3).astype(np.float32)
    for (alpha, beta) in [(1.0, 0.0), (1.5, 0.5)]:
        ref = np.asfortranarray(c0.copy()); oracle.gemm_f32(alpha, np.asfortranarray(a), np.asfortranarray(b), beta, ref)
        for order in ("F", "C"):
            got = c0.copy(order=order)
            if beta == 0.0:
                got[...] = np.nan
            nab.gemm_f32(alpha, a.copy(order=order), b.copy(order="F"), beta, got)
            tol = 4 * 70 * np.finfo(np.float32).eps * np.linalg.norm(a) * np.linalg.norm(b)
            assert np.abs(got - ref).max() <= tol


def test_gemm_full_size_linearity_on_device(L):
    """16384^3 (BASELINE configs[1]) through size-independent properties on the device:
    C(A, B1 + B2) = C(A, B1) + C(A, B2) and a column spot-check against a plain matvec."""
    import torch
    from nalgebra_b200 import _capi
    n = 16384
    dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    A = torch.empty(n * n, dtype=torch.float64, device=dev); B = torch.empty_like(A); Cd = torch.empty_like(A)
    _capi.check(L.na_fill_uniform_dev(A.data_ptr(), n, n, n, 1, s))
    _capi.check(L.na_fill_uniform_dev(B.data_ptr(), n, n, n, 2, s))
    _capi.check(L.na_dgemm_dev(n, n, n, 1.0, A.data_ptr(), 1, n, B.data_ptr(), 1, n, 0.0, Cd.data_ptr(), 1, n, s))
    Am, Bm, Cm = A.view(n, n).t(), B.view(n, n).t(), Cd.view(n, n).t()     # column-major buffers as torch views
    norm_a, norm_b = torch.linalg.norm(Am).item(), torch.linalg.norm(Bm).item()
    tol = 4 * n * EPS * norm_a * norm_b
    for j in (0, 777, n - 1):
        assert (Am @ Bm[:, j] - Cm[:, j]).abs().max().item() <= tol
    for i in (0, 4097, n - 1):
        assert (Am[i, :] @ Bm - Cm[i, :]).abs().max().item() <= tol
    # beta path at full size: C <- 0.5*A*B + 0.5*C must reproduce C
    C2 = Cd.clone()
    _capi.check(L.na_dgemm_dev(n, n, n, 0.5, A.data_ptr(), 1, n, B.data_ptr(), 1, n, 0.5, C2.data_ptr(), 1, n, s))
    assert (C2 - Cd).abs().max().item() <= tol
    # Regression (round 1): a shared-memory stage was handed back to the TMA producer while the last
    # fragment load of the stage was still in flight; under beta != 0 epilogues about one warp tile per
    # launch picked up 4 k-values of the wrong k-block (error ~1e-4 relative, far inside `tol`).  The kernel
    # is deterministic, so repeated launches must agree bit for bit over the whole matrix, and the beta path
    # must reproduce C to rounding (|C| ~ 4100, ulp ~ 9e-13).
    for _ in range(6):
        C3 = Cd.clone()
        _capi.check(L.na_dgemm_dev(n, n, n, 0.5, A.data_ptr(), 1, n, B.data_ptr(), 1, n, 0.5, C3.data_ptr(), 1, n, s))
        assert torch.equal(C3, C2)
        assert (C3 - Cd).abs().max().item() <= 1e-8
        del C3


def test_lu_panel_trsm_kernel(L):
    """The direct unit-lower TRSM of the LU panels (trsm_unit_lower_small_kernel) against numpy, through na_trsm_f64_dev: block sizes around 64/128, ragged right-hand-side counts, garbage on and above L's diagonal, padding rows
    untouched.  (Round 1: a __restrict__ shared-memory pointer let nvcc keep a stale value across __syncthreads.)"""
    import ctypes as C
    import torch
    from nalgebra_b200 import _capi
    def f(n1, l_ptr, ldl, b_ptr, ldb, nrhs, stream):      # na_trsm_f64_dev routes unit-lower solves of <= 128 rows to that kernel
        return L.na_trsm_f64_dev(0, 1, 0, 1, n1, nrhs, l_ptr, ldl, b_ptr, ldb, stream)
    s = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(1)
    for n1 in (1, 37, 64, 65, 100, 127, 128):
        for nrhs in (1, 3, 64, 65, 200):
            ldl, ldb = n1 + 6, n1 + 10
            lm = np.asfortranarray(np.tril(rng.random((ldl, n1)) - 0.5, -1))
            lfull = lm.copy(); lfull[:n1][np.triu_indices(n1)] = 7.7
            b = np.asfortranarray(rng.random((ldb, nrhs)))
            ref = np.linalg.solve(np.tril(lm[:n1], -1) + np.eye(n1), b[:n1])
            dl = torch.from_numpy(lfull.T.copy()).cuda(); db = torch.from_numpy(b.T.copy()).cuda()
            _capi.check(f(n1, dl.data_ptr(), ldl, db.data_ptr(), ldb, nrhs, s)); torch.cuda.synchronize()
            got = db.cpu().numpy().T
            assert np.abs(got[:n1] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), (n1, nrhs)
            assert np.array_equal(got[n1:], b[n1:]), (n1, nrhs)


# ---- Cholesky --------------------------------------------------------------------------------------
def test_cholesky_kats(nab):
    k = K["cholesky_with_substitute"]
    m = mat(k["m"])
    assert nab.Cholesky.new(m) is None
    ch = nab.Cholesky.new_with_substitute(m, k["substitute"])
    assert ch is not None and np.isnan(ch.l_dirty()[0, 1])      # strict upper never touched


@pytest.mark.parametrize("n", list(range(1, 21)) + [64, 128, 129, 257, 640, 1000])
def test_cholesky_vs_oracle(nab, oracle, n):
    rng = np.random.default_rng(n)
    m = random_sdp(n, rng) if n <= 20 else oracle.spd_wellcond(n, 5)
    ch = nab.Cholesky.new(m)
    assert ch is not None
    lref = np.tril(oracle.cholesky(m))
    l = ch.l()
    assert np.abs(l - lref).max() <= 10 * n * EPS * np.abs(lref).max()
    assert np.linalg.norm(l @ l.T - m) / np.linalg.norm(m) <= 10 * n * EPS
    assert relative_eq(m, l @ l.T, 1e-7)                          # the reference's own proptest bound
    b = rng.random((n, 3))
    x = ch.solve(b)
    assert relative_eq(m @ x, b, 1e-7)
    xr = oracle.cholesky_solve(oracle.cholesky(m), b)
    assert np.abs(x - xr).max() <= 1e-10 * max(1.0, np.abs(xr).max())
    if n <= 20:
        assert np.allclose(ch.inverse() @ m, np.eye(n), atol=1e-7)
        assert np.isclose(ch.determinant(), np.linalg.det(m), rtol=1e-9)


def test_cholesky_failure_column_and_bench_spd(nab, oracle):
    m = oracle.spd_wellcond(300, 5)
    m[150, 150] = -1.0
    assert nab.Cholesky.new(m) is None and oracle.cholesky(m) is None
    from nalgebra_b200 import _capi
    a = m.copy(order="F"); fail = C.c_size_t(0)
    st = _capi.lib().na_cholesky_f64(300, a.ctypes.data, 300, 0, 0.0, C.addressof(fail))
    assert st == _capi.NA_NOT_PD and fail.value == 150
    rng = np.random.default_rng(3)
    m = bench_spd(200, rng)                                       # the reference benches' ill-conditioned recipe
    ch = nab.Cholesky.new(m)
    assert ch is not None and np.linalg.norm(ch.l() @ ch.l().T - m) / np.linalg.norm(m) <= 10 * 200 * EPS


# ---- LU --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["lu_simple", "lu_simple_with_pivot"])
def test_lu_kats(nab, name):
    k = K[name]
    m = mat(k["m"])
    lu = nab.LU.new(m)
    assert lu.determinant() == k["determinant"]                  # exact, like the reference's assert_eq!
    p, l, u = lu.unpack()
    rec = l @ u; p.inv_permute_rows(rec)
    assert relative_eq(m, rec, k["epsilon"])


def test_matrix5_try_inverse(nab):
    k = K["matrix5_try_inverse"]
    inv = nab.LU.new(mat(k["a"])).try_inverse()
    exp = mat(k["expected_inverse"])
    d = np.abs(inv - exp)
    assert np.all((d <= k["max_relative"] * np.maximum(np.abs(inv), np.abs(exp))) | (np.abs(exp) < 1e-15))


@pytest.mark.parametrize("shape", [(n, n) for n in range(1, 21)] + [(3, 5), (5, 3), (4, 4), (64, 64), (128, 128), (129, 129),
                                   (130, 257), (257, 130), (300, 300), (640, 640), (1000, 1000), (1537, 1537)])
def test_lu_pivots_bit_exact_vs_oracle(nab, oracle, shape):
    m, n = shape
    a = oracle.uniform(m, n, 6) - 0.3
    lu = nab.LU.new(a)
    lur, swr = oracle.lu(a)
    assert np.array_equal(lu.p().ipiv, swr)                       # PermutationSequence bit-exact
    assert np.abs(lu.lu_internal() - lur).max() <= 1e-9
    p, l, u = lu.unpack()
    rec = l @ u; p.inv_permute_rows(rec)
    assert np.linalg.norm(rec - a) / np.linalg.norm(a) <= 10 * max(m, n) * EPS
    if m == n:
        b = oracle.uniform(n, 3, 7)
        x = lu.solve(b); xr = oracle.lu_solve(lur, swr, b)
        assert x is not None and np.abs(x - xr).max() <= 1e-7 * max(1.0, np.abs(xr).max())


def test_lu_2048_pivots_and_64_rhs(nab, oracle):
    """Scaled-down BASELINE configs[3]: pivots bit-exact with the CPU oracle, then solve with 64 RHS."""
    n = 2048
    a = oracle.uniform(n, n, 6)
    lu = nab.LU.new(a)
    lur, swr = oracle.lu(a)
    assert np.array_equal(lu.p().ipiv, swr)
    b = oracle.uniform(n, 64, 7)
    x = lu.solve(b)
    assert np.linalg.norm(a @ x - b) / (np.linalg.norm(a) * np.linalg.norm(x)) <= 10 * n * EPS


def test_lu_zero_column_singular_and_ties(nab, oracle):
    a = oracle.uniform(50, 50, 6)
    a[:, 3] = 0.0
    lu = nab.LU.new(a); lur, swr = oracle.lu(a)
    assert np.array_equal(lu.p().ipiv, swr) and np.abs(lu.lu_internal() - lur).max() < 1e-10
    assert lu.solve(np.ones((50, 1))) is None and not lu.is_invertible()
    t = np.asfortranarray(np.array([[1.0, 2.0, 0.0], [-3.0, 1.0, 1.0], [3.0, 0.0, 2.0], [2.0, 5.0, 1.0]]))   # |x| tie: lowest index wins
    lu = nab.LU.new(t); lur, swr = oracle.lu(t)
    assert np.array_equal(lu.p().ipiv, swr) and int(lu.p().ipiv[0, 1]) == 1
    assert np.array_equal(nab.LU.new(np.eye(5)).p().ipiv.shape, (0, 2))


# ---- QR --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(n, n) for n in range(1, 21)] + [(5, 3), (3, 5), (4, 4), (33, 33), (64, 20), (100, 64),
                                   (257, 130), (130, 257), (300, 300), (700, 300), (1000, 513), (2000, 600)])
def test_qr_vs_oracle(nab, oracle, shape):
    m, n = shape
    k = min(m, n)
    a = oracle.uniform(m, n, 8) - 0.5
    qr = nab.QR.new(a)
    qref, dref = oracle.qr(a)
    assert np.abs(qr.qr_internal() - qref).max() <= 1e-10        # nalgebra storage: unit axes + strict-upper R
    assert np.abs(qr.diag_internal() - dref).max() <= 1e-10      # signed diag
    q, r = qr.unpack()
    assert np.linalg.norm(q @ r - a) / np.linalg.norm(a) <= 10 * max(m, n) * EPS
    assert np.linalg.norm(q.T @ q - np.eye(k)) <= 10 * max(m, n) * EPS
    assert np.abs(q - oracle.qr_q(qref, dref)).max() <= 1e-10
    b = oracle.uniform(m, 3, 4)
    bq = b.copy(order="F"); qr.q_tr_mul(bq)
    assert np.abs(bq - oracle.qr_q_tr_mul(qref, dref, b)).max() <= 1e-10
    if m == n:
        x = qr.solve(b)
        assert x is not None and np.allclose(a @ x, b, rtol=1e-6, atol=1e-6)
        assert qr.is_invertible()


def test_qr_zero_column_and_singular(nab, oracle):
    a = oracle.uniform(20, 12, 8) - 0.5
    a[:, 4] = 0.0
    qr = nab.QR.new(a); qref, dref = oracle.qr(a)
    assert np.abs(qr.qr_internal() - qref).max() <= 1e-12 and np.abs(qr.diag_internal() - dref).max() <= 1e-12
    z = np.zeros((6, 6)); z[:, 1:] = oracle.uniform(6, 5, 1)
    qr = nab.QR.new(z)
    assert qr.diag_internal()[0] == 0.0 and not qr.is_invertible() and qr.solve(np.ones((6, 1))) is None


# ---- triangular solves ------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 4, 5, 100, 300])
def test_triangular_solves(nab, oracle, n):
    a = oracle.uniform(n, n, 9) + n * np.eye(n)
    b = oracle.uniform(n, 5, 10)
    assert np.allclose(nab.solve_lower_triangular(a, b), oracle.solve_lower(a, b), rtol=1e-12, atol=1e-13)
    assert np.allclose(nab.solve_upper_triangular(a, b), oracle.solve_upper(a, b), rtol=1e-12, atol=1e-13)
    assert np.allclose(np.tril(a).T @ nab.tr_solve_lower_triangular(a, b), b)
    assert np.allclose(np.triu(a).T @ nab.tr_solve_upper_triangular(a, b), b)
    x = nab.solve_lower_triangular_with_diag(a, b, 1.0)
    assert np.allclose((np.tril(a, -1) + np.eye(n)) @ x, b)
    if n > 2:
        a[1, 1] = 0.0
        assert nab.solve_lower_triangular(a, b) is None and nab.solve_upper_triangular(a, b) is None


# ---- BASELINE sizes through properties, on the device -------------------------------------------------
def test_cholesky_16384_residual(L):
    import torch
    from nalgebra_b200 import _capi
    n = 16384
    dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    A0 = torch.empty(n * n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), n, n, n, 5, s))
    M = A0.view(n, n); M.copy_((M + M.t()) * 0.5); M.diagonal().add_(float(n))
    A = A0.clone(); fail = C.c_size_t(0)
    assert _capi.check(L.na_cholesky_f64_dev(n, A.data_ptr(), n, 0, 0.0, C.addressof(fail), s)) == 0
    Lm = torch.tril(A.view(n, n).t())
    R = torch.tril(Lm @ Lm.t() - M.t())
    assert (torch.linalg.norm(R) / torch.linalg.norm(torch.tril(M.t()))).item() <= 10 * n * EPS
    assert torch.equal(torch.triu(A.view(n, n).t(), 1), torch.triu(M.t(), 1))     # strict upper untouched


def test_lu_16384_residual_and_64_rhs(L):
    import torch
    from nalgebra_b200 import _capi
    n, nrhs = 16384, 64
    dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    A0 = torch.empty(n * n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), n, n, n, 6, s))
    A = A0.clone(); swaps = (C.c_size_t * (2 * n))(); ns = C.c_size_t(0)
    _capi.check(L.na_lu_f64_dev(n, n, A.data_ptr(), n, swaps, C.addressof(ns), s))
    LU = A.view(n, n).t()
    Lm = torch.tril(LU, -1); Lm.diagonal().add_(1.0)
    sw = np.frombuffer(swaps, dtype=np.uint64)[: 2 * ns.value].reshape(-1, 2).astype(np.int64)
    assert np.all(sw[:, 1] > sw[:, 0]) and np.all(np.diff(sw[:, 0]) > 0)      # PermutationSequence layout
    perm = np.arange(n)
    for i, j in sw:
        perm[[i, j]] = perm[[j, i]]
    PA = A0.view(n, n).t()[torch.from_numpy(perm).to(dev)]
    assert (torch.linalg.norm(PA - Lm @ torch.triu(LU)) / torch.linalg.norm(PA)).item() <= 10 * n * EPS
    assert (Lm.abs().max().item() <= 1.0)                                       # partial pivoting: |l_ij| <= 1
    del Lm, PA
    B = torch.empty(n * nrhs, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(B.data_ptr(), n, nrhs, n, 7, s))
    X = B.clone()
    assert _capi.check(L.na_lu_solve_f64_dev(n, A.data_ptr(), n, swaps, ns.value, X.data_ptr(), n, nrhs, s)) == 0
    Am, Xm, Bm = A0.view(n, n).t(), X.view(nrhs, n).t(), B.view(nrhs, n).t()
    res = torch.linalg.norm(Am @ Xm - Bm) / (torch.linalg.norm(Am) * torch.linalg.norm(Xm))
    assert res.item() <= 10 * n * EPS


def test_qr_65536x4096_residual_and_orthogonality(L):
    import torch
    from nalgebra_b200 import _capi
    m, n = 65536, 4096
    dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    A0 = torch.empty(m * n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), m, n, m, 8, s))
    A = A0.clone(); dg = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_qr_f64_dev(m, n, A.data_ptr(), m, dg.data_ptr(), s))
    Q = torch.empty(m * n, dtype=torch.float64, device=dev)
    _capi.check(L.na_qr_q_f64_dev(m, n, A.data_ptr(), m, dg.data_ptr(), Q.data_ptr(), m, s))
    Qm = Q.view(n, m).t()
    R = torch.triu(A.view(n, m).t()[:n, :]).clone(); R.diagonal().copy_(dg.abs())
    A0m = A0.view(n, m).t()
    assert (torch.linalg.norm(A0m - Qm @ R) / torch.linalg.norm(A0m)).item() <= 10 * m * EPS
    assert torch.linalg.norm(Qm.t() @ Qm - torch.eye(n, device=dev, dtype=torch.float64)).item() <= 10 * m * EPS
    axes = torch.tril(A.view(n, m).t())
    assert (torch.linalg.norm(axes, dim=0) - 1.0).abs().max().item() <= 1e-12   # unit Householder axes


@pytest.mark.parametrize("reg_leaf,fused", [(1, 1), (0, 1), (0, 0)])
def test_qr_tall_paths_vs_oracle(nab, oracle, L, reg_leaf, fused):
    """Tall QR (m >= 8192, >= 4 outer panels) along every driver: register-resident leaf + plain loop (default),
    shared-memory leaf + two-stream look-ahead, and the GEMM-sequence in-panel reflectors, against the oracle."""
    from nalgebra_b200 import _capi
    m, n = 8200, 1030
    a = oracle.uniform(m, n, 8) - 0.4
    _capi.check(L.na_set_tuning(b"qr_reg_leaf", reg_leaf)); _capi.check(L.na_set_tuning(b"qr_fused", fused))
    try:
        qr = nab.QR.new(a)
    finally:
        _capi.check(L.na_set_tuning(b"qr_reg_leaf", 1)); _capi.check(L.na_set_tuning(b"qr_fused", 1))
    qr_ref, diag_ref = oracle.qr(a)
    assert np.abs(qr.qr_internal() - qr_ref).max() < 1e-10
    assert np.abs(qr.diag_internal() - diag_ref).max() < 1e-10
    q, r = qr.q(), qr.r()
    assert np.linalg.norm(q @ r - a) / np.linalg.norm(a) <= 10 * m * EPS
    assert np.abs(q.T @ q - np.eye(n)).max() <= 10 * m * EPS


@pytest.mark.parametrize("shape", [(2100, 70), (2500, 300), (3000, 333), (4100, 257), (5000, 40)])
def test_qr_fused_reflector_ragged_shapes(nab, oracle, shape):
    """The fused in-panel block reflector (leaves with >= 2048 rows) and the register-resident leaf (>= 4096 rows) on
    widths that are not multiples of 32 or 8, with a zero column (tau = 0) inside a fused leaf."""
    m, n = shape
    a = oracle.uniform(m, n, 8) - 0.4
    if n > 45: a[:, 40] = 0.0
    qr = nab.QR.new(a)
    qr_ref, diag_ref = oracle.qr(a)
    assert np.abs(qr.qr_internal() - qr_ref).max() < 1e-10
    assert np.abs(qr.diag_internal() - diag_ref).max() < 1e-10


def test_host_pointer_calls_stream_and_match_the_device_calls(nab, L):
    """na_cholesky_f64 (lower triangle only over PCIe, finished panels streamed back) and na_qr_f64 (panels converted and
    streamed back behind the factorization) must give the bits of the device-resident calls; Cholesky must leave the
    host's strict upper triangle alone and report the failing column as before."""
    import ctypes as C
    import torch
    from nalgebra_b200 import _capi
    dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    n = 2304                                                   # > 2 * 512: the look-ahead / streaming path
    a0 = torch.empty(n * n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_spd_block_dev(a0.data_ptr(), n, n, n, 5, 0, 0, n, s))
    a = a0.clone(); fail = C.c_size_t(0)
    assert _capi.check(L.na_cholesky_f64_dev(n, a.data_ptr(), n, 0, 0.0, C.addressof(fail), s)) == 0
    h0 = a0.cpu(); h = h0.clone().pin_memory()
    assert _capi.check(L.na_cholesky_f64(n, h.data_ptr(), n, 0, 0.0, C.addressof(fail))) == 0
    g, hm = a.cpu().view(n, n).t(), h.view(n, n).t()
    assert torch.equal(torch.tril(g), torch.tril(hm))
    assert torch.equal(torch.triu(hm, 1), torch.triu(h0.view(n, n).t(), 1))
    h.copy_(h0); h.view(n, n)[1500, 1500] = -1.0
    assert L.na_cholesky_f64(n, h.data_ptr(), n, 0, 0.0, C.addressof(fail)) == 1 and fail.value == 1500
    m, nq = 8300, 1100
    a0 = torch.empty(m * nq, dtype=torch.float64, device=dev); d = torch.empty(nq, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(a0.data_ptr(), m, nq, m, 8, s))
    a = a0.clone()
    _capi.check(L.na_qr_f64_dev(m, nq, a.data_ptr(), m, d.data_ptr(), s)); torch.cuda.synchronize()
    h = a0.cpu().pin_memory(); hd = torch.empty(nq, dtype=torch.float64)
    _capi.check(L.na_qr_f64(m, nq, h.data_ptr(), m, hd.data_ptr()))
    assert torch.equal(a.cpu(), h) and torch.equal(d.cpu(), hd)


def test_concurrent_calls_from_two_host_threads(L):
    """The C ABI is called concurrently from two host threads on their own streams (nalgebra's types are Send + Sync,
    SURVEY 8b 'Threading'): per-call workspaces, thread-local SM limits.  Results must equal the serial ones bit for bit."""
    import ctypes as C
    import threading
    import torch
    from nalgebra_b200 import _capi
    n = 2560
    dev = torch.device("cuda:0")
    s0 = torch.cuda.current_stream().cuda_stream
    base_lu = torch.empty(n * n, dtype=torch.float64, device=dev); base_ch = torch.empty_like(base_lu)
    _capi.check(L.na_fill_uniform_dev(base_lu.data_ptr(), n, n, n, 6, s0))
    _capi.check(L.na_fill_spd_block_dev(base_ch.data_ptr(), n, n, n, 5, 0, 0, n, s0))
    torch.cuda.synchronize()

    def run_lu(stream, out):
        a = base_lu.clone(); torch.cuda.synchronize()
        swaps = (C.c_size_t * (2 * n))(); ns = C.c_size_t(0)
        _capi.check(L.na_lu_f64_dev(n, n, a.data_ptr(), n, swaps, C.addressof(ns), stream))
        torch.cuda.synchronize()
        out["lu"] = (a.cpu(), list(swaps[: 2 * ns.value]))

    def run_chol(stream, out):
        a = base_ch.clone(); torch.cuda.synchronize()
        fail = C.c_size_t(0)
        _capi.check(L.na_cholesky_f64_dev(n, a.data_ptr(), n, 0, 0.0, C.addressof(fail), stream))
        torch.cuda.synchronize()
        out["chol"] = a.cpu()

    serial = {}
    run_lu(s0, serial); run_chol(s0, serial)
    for _ in range(3):
        st1, st2 = torch.cuda.Stream(), torch.cuda.Stream()
        conc = {}
        t1 = threading.Thread(target=run_lu, args=(st1.cuda_stream, conc))
        t2 = threading.Thread(target=run_chol, args=(st2.cuda_stream, conc))
        t1.start(); t2.start(); t1.join(); t2.join()
        assert conc["lu"][1] == serial["lu"][1]
        assert torch.equal(conc["lu"][0], serial["lu"][0])
        assert torch.equal(conc["chol"], serial["chol"])



# ---- round 2: parity holes the round-1 review named ---------------------------------------------------
def test_vector_right_hand_sides(nab, oracle):
    """A DVector rhs (1-D array) is an n x 1 matrix: solve / solve_mut / q_tr_mul / triangular solves (round 1 turned it
    into 1 x n and read n*n doubles from an n-double buffer)."""
    n = 37
    spd = oracle.spd_wellcond(n, 5); a = oracle.uniform(n, n, 6) - 0.3
    b = oracle.uniform(n, 1, 7)[:, 0].copy()
    ch = nab.Cholesky.new(spd); x = ch.solve(b)
    assert x.shape == (n,) and np.abs(spd @ x - b).max() <= 1e-10
    b2 = b.copy(); ch.solve_mut(b2); assert np.array_equal(b2, x)
    lu = nab.LU.new(a); x = lu.solve(b)
    assert x.shape == (n,) and np.abs(a @ x - b).max() <= 1e-9
    qr = nab.QR.new(a); x = qr.solve(b)
    assert x.shape == (n,) and np.abs(a @ x - b).max() <= 1e-9
    v = b.copy(); qr.q_tr_mul(v)
    assert np.abs(v - qr.q().T @ b).max() <= 1e-12
    t = np.tril(a) + 3 * np.eye(n)
    x = nab.solve_lower_triangular(t, b)
    assert x.shape == (n,) and np.abs(t @ x - b).max() <= 1e-10
    with pytest.raises(ValueError):
        ch.solve(np.ones(n + 1))


def test_solve_lower_triangular_with_diag(nab, oracle):
    """solve.rs:106-133 including its quirk: the diagonal is `diag`, and b[i] is never divided by it."""
    C_ = C
    for n, diag in ((1, 2.5), (7, 1.0), (33, -0.75), (130, 3.0)):
        t = np.asfortranarray(oracle.uniform(n, n, 3) - 0.5); b = np.asfortranarray(oracle.uniform(n, 3, 4))
        ref = b.copy(order="F")
        assert oracle.lib().na_oracle_solve_lower_with_diag_f64(n, t.ctypes.data, n, C_.c_double(diag), ref.ctypes.data, n, 3) == 1
        got = nab.solve_lower_triangular_with_diag(t, b, diag)
        assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), (n, diag)
    assert nab.solve_lower_triangular_with_diag(np.eye(3), np.ones(3), 0.0) is None


def test_api_names_of_the_reference(nab, oracle):
    """tr_mul_to / ad_mul(_to) / gemm_ad (ops.rs:674-779, blas.rs:827-860), LU::try_inverse_to / l_unpack,
    lu::try_invert_to (lu.rs:51-86, 156-171, 283-297)."""
    a = oracle.uniform(40, 23, 1) - 0.5; b = oracle.uniform(40, 17, 2) - 0.5
    out = np.full((23, 17), np.nan, order="F")
    nab.tr_mul_to(a, b, out)
    assert np.abs(out - a.T @ b).max() <= gemm_tol(a, b, 40)
    assert np.array_equal(nab.ad_mul(a, b), nab.tr_mul(a, b))
    with pytest.raises(ValueError):
        nab.tr_mul_to(a, b, np.empty((17, 23), order="F"))
    m = oracle.uniform(30, 30, 6) - 0.3
    inv = np.empty((30, 30), order="F")
    assert nab.try_invert_to(m, inv) and np.abs(inv @ m - np.eye(30)).max() <= 1e-9
    inv2 = np.empty((30, 30), order="F")
    lu = nab.LU.new(m)
    assert lu.try_inverse_to(inv2) and np.array_equal(inv, inv2)
    l = lu.l(); assert np.array_equal(lu.l_unpack(), l)
    sing = m.copy(); sing[:, 4] = 0.0
    assert not nab.try_invert_to(sing, inv)


def test_syrk_lower_host_entry(nab, oracle):
    """na_dsyrk_lower: C <- alpha*A*A^T + beta*C on the lower triangle only; the strict upper triangle of the host matrix
    (NaN here) is neither read nor written.  The reference's SPD recipes form M*M^T (benches/linalg/cholesky.rs:3-11)."""
    for (n, k) in ((1, 1), (5, 3), (130, 70), (257, 300), (600, 64)):
        a = oracle.uniform(n, k, 4) - 0.5
        for alpha, beta in ((1.0, 0.0), (1.5, 0.5)):
            c0 = oracle.uniform(n, n, 9)
            c = np.asfortranarray(np.tril(c0) + np.triu(np.full((n, n), np.nan), 1))
            nab.syrk_lower(alpha, a, beta, c)
            ref = alpha * (a @ a.T) + (beta * c0 if beta else 0.0)
            assert np.all(np.isnan(c[np.triu_indices(n, 1)])), (n, k)
            assert np.abs(np.tril(c) - np.tril(ref)).max() <= gemm_tol(a, a, k) + 1e-14, (n, k, alpha, beta)
    # the bench SPD recipe end to end: M*M^T + sqrt(eps)*|M|_F^2*I on the GPU, then Cholesky::new of it
    n = 300
    m = oracle.uniform(n, n, 4)
    c = np.zeros((n, n), order="F")
    nab.syrk_lower(1.0, m, 0.0, c)
    c[np.diag_indices(n)] += np.sqrt(EPS) * np.linalg.norm(m) ** 2
    assert nab.Cholesky.new(c) is not None
    # k == 0: the beta contract of gemm_uninit (blas_uninit.rs:258-269) on the triangle
    c = np.asfortranarray(np.tril(np.ones((5, 5))) + np.triu(np.full((5, 5), np.nan), 1))
    nab.syrk_lower(1.0, np.zeros((5, 0)), 0.5, c)
    assert np.array_equal(np.tril(c), 0.5 * np.tril(np.ones((5, 5)))) and np.all(np.isnan(c[np.triu_indices(5, 1)]))


def test_gemv_and_axcpy(nab, oracle, L):
    """gemv_uninit / gemv_tr / axcpy (blas_uninit.rs:86-177, blas.rs:503-540): G8, the Level-1/2 fallbacks."""
    import torch
    from nalgebra_b200 import _capi
    for (m, n) in ((1, 1), (5, 3), (3, 5), (130, 70), (1000, 513), (2000, 3000)):
        a = oracle.uniform(m, n, 1) - 0.5; x = oracle.uniform(n, 1, 2)[:, 0] - 0.5; y0 = oracle.uniform(m, 1, 3)[:, 0]
        tol = 4 * n * EPS * np.linalg.norm(a) * np.linalg.norm(x) + 1e-300
        for alpha, beta in ((1.0, 0.0), (1.5, 0.5)):
            y = np.full(m, np.nan) if beta == 0.0 else y0.copy()
            nab.gemv(alpha, a, x, beta, y)
            assert np.abs(y - (alpha * a @ x + (beta * y0 if beta else 0.0))).max() <= tol, (m, n)
            # row-major A (transposed view) and a strided x / y
            y = np.zeros(2 * m); y[::2] = y0
            nab.gemv(alpha, np.ascontiguousarray(a), np.repeat(x, 2)[::2], beta, y[::2])
            assert np.abs(y[::2] - (alpha * a @ x + (beta * y0 if beta else 0.0))).max() <= tol
        yt = np.full(n, np.nan)
        nab.gemv_tr(1.0, a, y0, 0.0, yt)
        assert np.abs(yt - a.T @ y0).max() <= 4 * m * EPS * np.linalg.norm(a) * np.linalg.norm(y0) + 1e-300
    y = np.ones(4); nab.gemv(1.0, np.zeros((4, 0)), np.zeros(0), 0.5, y); assert np.array_equal(y, 0.5 * np.ones(4))
    # axcpy on the device: (a*x)*c + b*y with the reference's unfused rounding == the oracle's axcpy bit for bit
    n = 1000
    xs = oracle.uniform(n, 1, 5)[:, 0]; ys = oracle.uniform(n, 1, 6)[:, 0]
    dx = torch.from_numpy(xs).cuda(); dy = torch.from_numpy(ys.copy()).cuda()
    _capi.check(L.na_daxcpy_dev(n, 1.7, dx.data_ptr(), 1, 0.3, -2.5, dy.data_ptr(), 1, torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(dy.cpu().numpy(), (1.7 * xs) * 0.3 + (-2.5) * ys)


def test_lu_pivot_divergence_report(nab, oracle):
    """The tie-aware comparator (SURVEY "hard part 2"): equal sequences -> None; a forced divergence is located and the
    relative gap of the two candidates is what was planted."""
    from helpers import first_pivot_divergence
    n = 200
    a = oracle.uniform(n, n, 6) - 0.3
    lu = nab.LU.new(a)
    lur, swr = oracle.lu(a)
    assert first_pivot_divergence(a, lu.lu_internal(), lu.p().ipiv, swr) is None
    # plant a near-tie in column 0: rows r1 < r2 hold the two largest |values|, r2 larger by 1e-13 relative
    b = a.copy()
    col = np.abs(b[:, 0]); r1 = int(np.argmax(col)); big = col[r1]
    r2 = (r1 + 7) % n
    b[r2, 0] = big * (1 + 4e-13)
    lub = nab.LU.new(b)
    fake = lub.p().ipiv.copy()
    fake[0] = (0, r1) if r1 != 0 else fake[0]
    rep = first_pivot_divergence(b, lub.lu_internal(), lub.p().ipiv, fake)
    if r1 != 0 and r2 != 0:
        assert rep is not None and rep["step"] == 0 and rep["rows"] == (r2, r1) and 1e-13 < rep["relative_gap"] < 1e-12


def test_lu_8192_pivots_bit_exact_vs_oracle(nab, oracle, L):
    """BASELINE configs[3] asks for bit-exact pivots vs the CPU: the look-ahead driver (flat 512-panels, register-resident
    GETF2 leaves, DMMA trailing updates) against the oracle's unblocked, unfused LU at N = 8192 (~3 min of one CPU core)."""
    from helpers import first_pivot_divergence
    n = 8192
    a = oracle.uniform(n, n, 6)
    lu = nab.LU.new(a)
    lur, swr = oracle.lu(a)
    rep = first_pivot_divergence(a, lu.lu_internal(), lu.p().ipiv, swr)
    assert rep is None, rep
    assert np.abs(lu.lu_internal() - lur).max() <= 1e-8


def test_lu_16384_lookahead_and_plain_paths_agree(L):
    """N = 16384: the two-stream look-ahead driver and the plain recursive driver (different blocking of the trailing
    updates) must produce the same PermutationSequence; the factors agree to rounding."""
    import torch
    from nalgebra_b200 import _capi
    n = 16384
    dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    A0 = torch.empty(n * n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), n, n, n, 6, s))
    res = []
    try:
        for la in (1, 0):
            _capi.check(L.na_set_tuning(b"lu_lookahead", la))
            A = A0.clone(); swaps = (C.c_size_t * (2 * n))(); ns = C.c_size_t(0)
            _capi.check(L.na_lu_f64_dev(n, n, A.data_ptr(), n, swaps, C.addressof(ns), s))
            res.append((A, np.frombuffer(swaps, dtype=np.uint64)[: 2 * ns.value].copy()))
    finally:
        _capi.check(L.na_set_tuning(b"lu_lookahead", 1))
    if not np.array_equal(res[0][1], res[1][1]):
        from helpers import first_pivot_divergence
        rep = first_pivot_divergence(A0.view(n, n).t().cpu().numpy(), res[0][0].view(n, n).t().cpu().numpy(),
                                     res[0][1].reshape(-1, 2), res[1][1].reshape(-1, 2))
        pytest.fail(f"pivot sequences differ: {rep}")
    assert (res[0][0] - res[1][0]).abs().max().item() <= 1e-7


# ---- the LAPACK-symbol facade (nalgebra-lapack --features lapack-custom) against scipy's LAPACK ----------------------
def _f(lib, name, *args):
    """Fortran ABI: every argument by pointer."""
    keep, ptrs = [], []
    for a in args:
        if isinstance(a, np.ndarray):
            ptrs.append(C.c_void_p(a.ctypes.data)); keep.append(a)
        elif isinstance(a, bytes):
            b = C.create_string_buffer(a); keep.append(b); ptrs.append(C.cast(b, C.c_void_p))
        elif isinstance(a, int):
            v = C.c_int(a); keep.append(v); ptrs.append(C.cast(C.byref(v), C.c_void_p))
        else:
            raise TypeError(type(a))
    getattr(lib, name)(*ptrs)
    return keep


def test_lapack_facade_vs_scipy(L, oracle):
    """dpotrf_/dpotrs_/dpotri_, dgetrf_/dgetrs_/dgetri_/dlaswp_, dgeqrf_/dorgqr_/dormqr_, dtrtrs_ with the Fortran ABI and
    LAPACK layouts, as nalgebra-lapack calls them (nalgebra-lapack/src/cholesky.rs:181-224, lu.rs:351-446, qr.rs:166-590)."""
    import scipy.linalg.lapack as sl
    for n in (1, 7, 130, 600):
        spd = oracle.spd_wellcond(n, 5)
        for uplo in (b"L", b"U"):
            a = np.asfortranarray(spd.copy()); info = np.zeros(1, dtype=np.int32)
            other = np.triu(a, 1) if uplo == b"L" else np.tril(a, -1)
            _f(L, "dpotrf_", uplo, n, a, n, info)
            ref, _ = sl.dpotrf(spd, lower=(uplo == b"L"), clean=False)
            tri = np.tril if uplo == b"L" else np.triu
            assert info[0] == 0 and np.abs(tri(a) - tri(ref)).max() <= 1e-11 * np.abs(ref).max(), (n, uplo)
            assert np.array_equal(np.triu(a, 1) if uplo == b"L" else np.tril(a, -1), other)          # other triangle untouched
            b = np.asfortranarray(oracle.uniform(n, 3, 7)); x = b.copy(order="F")
            _f(L, "dpotrs_", uplo, n, 3, a, n, x, n, info)
            assert info[0] == 0 and np.abs(spd @ x - b).max() <= 1e-9
            inv = a.copy(order="F")
            _f(L, "dpotri_", uplo, n, inv, n, info)
            assert info[0] == 0 and np.abs(tri(inv) - tri(np.linalg.inv(spd))).max() <= 1e-10
            assert np.array_equal(np.triu(inv, 1) if uplo == b"L" else np.tril(inv, -1), other)
    bad = oracle.spd_wellcond(50, 5); bad[20, 20] = -1.0
    a = np.asfortranarray(bad); info = np.zeros(1, dtype=np.int32)
    _f(L, "dpotrf_", b"L", 50, a, 50, info)
    assert info[0] == 21                                                       # leading minor of order 21 not positive definite

    for (m, n) in ((1, 1), (5, 3), (3, 5), (130, 130), (700, 700), (300, 520)):
        a0 = oracle.uniform(m, n, 6) - 0.3
        a = np.asfortranarray(a0.copy()); ipiv = np.zeros(min(m, n), dtype=np.int32); info = np.zeros(1, dtype=np.int32)
        _f(L, "dgetrf_", m, n, a, m, ipiv, info)
        lu_ref, piv_ref, _ = sl.dgetrf(a0)
        assert info[0] == 0 and np.array_equal(ipiv - 1, piv_ref), (m, n)           # same pivots as LAPACK (1-based on our side)
        assert np.abs(a - lu_ref).max() <= 1e-10, (m, n)
        if m == n:
            b = np.asfortranarray(oracle.uniform(n, 4, 7))
            for trans in (b"N", b"T"):
                x = b.copy(order="F")
                _f(L, "dgetrs_", trans, n, 4, a, n, ipiv, x, n, info)
                op = a0 if trans == b"N" else a0.T
                assert info[0] == 0 and np.abs(op @ x - b).max() <= 1e-8, (n, trans)
            inv = a.copy(order="F"); work = np.zeros(1); 
            _f(L, "dgetri_", n, inv, n, ipiv, work, -1, info); assert work[0] >= 1 and info[0] == 0       # workspace query
            _f(L, "dgetri_", n, inv, n, ipiv, np.zeros(max(n, 1)), n, info)
            assert info[0] == 0 and np.abs(inv @ a0 - np.eye(n)).max() <= 1e-8
            c = np.asfortranarray(oracle.uniform(n, 5, 9)); c2 = c.copy(order="F")
            _f(L, "dlaswp_", 5, c2, n, 1, n, ipiv, 1)
            assert np.array_equal(c2, sl.dlaswp(c, piv_ref, k1=0, k2=n - 1))
    sing = oracle.uniform(40, 40, 6); sing[:, 7] = 0.0
    a = np.asfortranarray(sing); ipiv = np.zeros(40, dtype=np.int32); info = np.zeros(1, dtype=np.int32)
    _f(L, "dgetrf_", 40, 40, a, 40, ipiv, info)
    assert info[0] == 8                                                        # U(8,8) exactly zero

    for (m, n) in ((1, 1), (7, 5), (130, 64), (600, 300), (300, 300), (2000, 260)):
        a0 = oracle.uniform(m, n, 8) - 0.5
        k = min(m, n)
        a = np.asfortranarray(a0.copy()); tau = np.zeros(k); info = np.zeros(1, dtype=np.int32); work = np.zeros(1)
        _f(L, "dgeqrf_", m, n, a, m, tau, work, -1, info); assert work[0] >= 1
        _f(L, "dgeqrf_", m, n, a, m, tau, work, 1, info)
        qr_ref, tau_ref, _, _ = sl.dgeqrf(a0)
        assert info[0] == 0
        # R equals LAPACK's up to the sign of a row: a column with nothing below its head (the last one of a square
        # matrix) is skipped by dlarfg (tau = 0) but reflected here, as nalgebra does (householder.rs:19-53)
        assert np.abs(np.abs(np.triu(a[:k])) - np.abs(np.triu(qr_ref[:k]))).max() <= 1e-10 * max(1.0, np.abs(qr_ref).max()), (m, n)
        full = slice(0, k - 1) if m == n else slice(0, k)
        if tau[full].size:
            assert np.abs(tau[full] - tau_ref[full]).max() <= 1e-11
            assert np.abs(np.tril(a, -1)[:, full] - np.tril(qr_ref, -1)[:, full]).max() <= 1e-10, (m, n)   # same reflector vectors
        q = np.asfortranarray(a[:, :k].copy())
        _f(L, "dorgqr_", m, k, k, q, m, tau, work, 1, info)
        assert info[0] == 0 and np.abs(q.T @ q - np.eye(k)).max() <= 10 * m * EPS
        assert np.abs(q @ np.triu(a[:k]) - a0).max() <= 10 * m * EPS * max(1.0, np.abs(a0).max()) * np.sqrt(n)
        cmat = np.asfortranarray(oracle.uniform(m, 3, 4))
        c1 = cmat.copy(order="F")
        _f(L, "dormqr_", b"L", b"T", m, 3, k, a, m, tau, c1, m, work, 1, info)       # Q^T C: its first k rows are Q_k^T C
        assert info[0] == 0 and np.abs(c1[:k] - q.T @ cmat).max() <= 1e-10
        assert abs(np.linalg.norm(c1) - np.linalg.norm(cmat)) <= 1e-10 * np.linalg.norm(cmat)
        _f(L, "dormqr_", b"L", b"N", m, 3, k, a, m, tau, c1, m, work, 1, info)       # Q (Q^T C) = C
        assert info[0] == 0 and np.abs(c1 - cmat).max() <= 1e-10
        cr = np.asfortranarray(oracle.uniform(4, m, 5)); c2 = cr.copy(order="F")
        _f(L, "dormqr_", b"R", b"N", 4, m, k, a, m, tau, c2, 4, work, 1, info)       # C Q: its first k columns are C Q_k
        assert info[0] == 0 and np.abs(c2[:, :k] - cr @ q).max() <= 1e-10
        _f(L, "dormqr_", b"R", b"T", 4, m, k, a, m, tau, c2, 4, work, 1, info)
        assert info[0] == 0 and np.abs(c2 - cr).max() <= 1e-10

    n = 200
    t = np.asfortranarray(oracle.uniform(n, n, 3) - 0.5 + 3 * np.eye(n)); b = np.asfortranarray(oracle.uniform(n, 3, 4))
    for uplo in (b"L", b"U"):
        for trans in (b"N", b"T"):
            for diag in (b"N", b"U"):
                x = b.copy(order="F"); info = np.zeros(1, dtype=np.int32)
                _f(L, "dtrtrs_", uplo, trans, diag, n, 3, t, n, x, n, info)
                ref, _ = sl.dtrtrs(t, b, lower=(uplo == b"L"), trans=(1 if trans == b"T" else 0), unitdiag=(diag == b"U"))
                assert info[0] == 0 and np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (uplo, trans, diag)
    t2 = t.copy(order="F"); t2[5, 5] = 0.0; x = b.copy(order="F"); info = np.zeros(1, dtype=np.int32)
    _f(L, "dtrtrs_", b"L", b"N", b"N", n, 3, t2, n, x, n, info)
    assert info[0] == 6 and np.array_equal(x, b)


# ---- f32 GEMM on tcgen05 (kind::tf32, 3xTF32, TMEM accumulators) ---------------------------------------------------------
EPS32 = float(np.finfo(np.float32).eps)


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 4), (7, 5, 3), (64, 64, 64), (65, 130, 31), (128, 128, 128), (129, 257, 100),
                                   (300, 200, 150), (513, 1000, 257), (1024, 1024, 1024)])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (1.5, 0.5)])
def test_sgemm_tc_vs_f64_all_layouts(nab, oracle, shape, alpha, beta):
    """na_sgemm (matrixmultiply::sgemm's seam, blas_uninit.rs:276-291) over the layout matrix of the f64 test: every
    combination of column-/row-major A, B, C.  Gate: |C - C_f64| <= 4*k*eps32*|A||B| (north_star, f32); the 3xTF32 split
    must in fact deliver f32-class accuracy, so a 100x tighter bound is asserted as well."""
    m, k, n = shape
    a64 = oracle.uniform(m, k, 1) - 0.5; b64 = oracle.uniform(k, n, 2) - 0.5; c64 = oracle.uniform(m, n, 3) - 0.5
    a32, b32, c32 = a64.astype(np.float32), b64.astype(np.float32), c64.astype(np.float32)
    ref = alpha * (a32.astype(np.float64) @ b32.astype(np.float64)) + (beta * c32.astype(np.float64) if beta else 0.0)
    tol = 4 * k * EPS32 * np.linalg.norm(a32) * np.linalg.norm(b32) + 1e-30
    for la in "FC":
        for lb in "FC":
            for lc in "FC":
                a = np.array(a32, order=la); b = np.array(b32, order=lb)
                c = np.array(c32 if beta else np.full((m, n), np.nan, dtype=np.float32), order=lc)
                nab.gemm_f32(alpha, a, b, beta, c)
                err = np.abs(c.astype(np.float64) - ref).max()
                assert err <= tol, (shape, la, lb, lc, err, tol)
                assert err <= 0.01 * tol + 8 * EPS32 * np.abs(ref).max(), (shape, la, lb, lc, err)


def test_sgemm_tc_views_and_k_zero(nab, oracle):
    a = (oracle.uniform(200, 300, 1) - 0.5).astype(np.float32); b = (oracle.uniform(300, 150, 2) - 0.5).astype(np.float32)
    # strided views (every other row / column), transposed operand, negative stride
    av, bv = a[::2, 1::2], b[1::2, ::3]
    c = np.zeros((100, 50), dtype=np.float32, order="F")
    nab.gemm_f32(1.0, av, bv, 0.0, c)
    assert np.abs(c - av.astype(np.float64) @ bv.astype(np.float64)).max() <= 1e-4
    c2 = np.zeros((100, 50), dtype=np.float32, order="F")
    nab.gemm_f32(1.0, av[::-1], bv, 0.0, c2)
    assert np.array_equal(c2, c[::-1])
    c3 = np.ones((4, 5), dtype=np.float32, order="F")
    nab.gemm_f32(1.0, np.zeros((4, 0), dtype=np.float32), np.zeros((0, 5), dtype=np.float32), 0.5, c3)
    assert np.array_equal(c3, np.full((4, 5), 0.5, dtype=np.float32))


def test_sgemm_tc_8192_device(L):
    """8192^3 on the device: linearity C(2A, B) == 2 C(A, B) bit for bit (scaling by 2 is exact in every stage of the 3xTF32
    pipeline), and a sampled comparison against float64."""
    import torch
    from nalgebra_b200 import _capi
    n = 8192
    dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    A = torch.rand(n * n, dtype=torch.float32, device=dev) - 0.5; B = torch.rand(n * n, dtype=torch.float32, device=dev) - 0.5
    Cd = torch.empty(n * n, dtype=torch.float32, device=dev); C2 = torch.empty_like(Cd)
    _capi.check(L.na_sgemm_dev(n, n, n, 1.0, A.data_ptr(), 1, n, B.data_ptr(), 1, n, 0.0, Cd.data_ptr(), 1, n, s))
    A2 = A * 2
    _capi.check(L.na_sgemm_dev(n, n, n, 1.0, A2.data_ptr(), 1, n, B.data_ptr(), 1, n, 0.0, C2.data_ptr(), 1, n, s))
    assert torch.equal(C2, Cd * 2)
    rows = torch.arange(0, n, 997, device=dev)
    ref = (A.view(n, n).t()[rows].double() @ B.view(n, n).t().double())          # column-major buffers viewed transposed
    got = Cd.view(n, n).t()[rows].double()
    assert (got - ref).abs().max().item() <= 3e-4                                 # f32-class: K accumulated in 256-wide TMEM chunks, summed in registers
