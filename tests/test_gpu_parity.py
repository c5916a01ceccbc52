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


def test_sgemm_vs_oracle(nab, oracle):
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
    """The direct unit-lower TRSM of the LU panel recursion (trsm_unit_lower_small_kernel) against numpy, through its
    test hook: block sizes around 64/128, ragged right-hand-side counts, garbage on and above L's diagonal, padding rows
    untouched.  (Round 1: a __restrict__ shared-memory pointer let nvcc keep a stale value across __syncthreads.)"""
    import ctypes as C
    import torch
    from nalgebra_b200 import _capi
    f = L.na_debug_trsm_unit_lower_small
    f.restype = C.c_int
    f.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
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


def test_qr_lookahead_path_vs_oracle(nab, oracle):
    """The two-stream look-ahead QR driver (m >= 8192, >= 4 outer panels) against the oracle's nalgebra restatement."""
    m, n = 8200, 1030
    a = oracle.uniform(m, n, 8) - 0.4
    qr = nab.QR.new(a)
    qr_ref, diag_ref = oracle.qr(a)
    assert np.abs(qr.qr_internal() - qr_ref).max() < 1e-10
    assert np.abs(qr.diag_internal() - diag_ref).max() < 1e-10
    q, r = qr.q(), qr.r()
    assert np.linalg.norm(q @ r - a) / np.linalg.norm(a) <= 10 * m * EPS
    assert np.abs(q.T @ q - np.eye(n)).max() <= 10 * m * EPS


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

