"""Hessenberg, SymmetricTridiagonal and Bidiagonal: the oracle pinned on the reference's own tests
(/root/reference/tests/linalg/hessenberg.rs:5-11 + proptests, tests/linalg/tridiagonal.rs:12-31,
tests/linalg/bidiagonal.rs:14-104: identity round trips are assert_eq!, the issue-1313 regressions) on the CPU; the CUDA
path against the oracle through the C ABI under -m gpu.  The kernels apply the reference's per-element update arithmetic
but sum the matrix-vector products in another order, so packed storage is compared to 1e-10 (1e-8 beyond n = 130; inputs of magnitude <= 1)
and the factorizations are checked through their reconstructions and through invariants (eigenvalues / singular values of
the reduced matrix)."""
import numpy as np
import pytest

from helpers import EPS, relative_eq

SQ = [1, 2, 3, 4, 5, 6, 7, 10, 13, 20, 33, 64, 100, 130]
RECT = [(1, 1), (2, 2), (5, 3), (3, 5), (4, 4), (7, 1), (1, 7), (10, 15), (15, 10), (20, 20), (33, 64), (64, 33), (130, 100)]


def _storage_tol(n):
    """Packed storage against the oracle: the entries are sums of O(n) products accumulated in another order, and the axes
    of the last few columns are ill-conditioned functions of the (by then tiny) trailing block: 1e-10 up to n = 130, 1e-8
    beyond.  The reconstruction residuals below are held to 10 n eps at every size."""
    return 1e-10 if n <= 130 else 1e-8


def _tri_matrix(diag, off):
    n = len(diag)
    t = np.diag(diag)
    if n > 1:
        t += np.diag(off, 1) + np.diag(off, -1)
    return t


# ---- CPU: the oracle against the reference's tests ---------------------------------------------------------------
def test_oracle_hessenberg_simple(oracle):
    m = np.asfortranarray(np.array([[1.0, 0.0], [1.0, 3.0]]))                  # hessenberg.rs:5-11
    hess, sub = oracle.hessenberg(m)
    p, h = oracle.assemble_q(hess, sub), oracle.hessenberg_h(hess, sub)
    assert relative_eq(m, p @ h @ p.T, 1e-7)


@pytest.mark.parametrize("n", SQ)
def test_oracle_hessenberg_and_tridiagonal_properties(oracle, n):
    a = oracle.uniform(n, n, 21) * 200.0 - 100.0                               # PROPTEST_F64 range
    hess, sub = oracle.hessenberg(a)
    p, h = oracle.assemble_q(hess, sub), oracle.hessenberg_h(hess, sub)
    assert relative_eq(a, p @ h @ p.T, 1e-7 * max(1.0, np.abs(a).max()))
    assert np.abs(p.T @ p - np.eye(n)).max() <= 1e-12
    assert np.abs(np.tril(h, -2)).max(initial=0.0) == 0.0
    if n > 2:                                                                   # second opinion: LAPACK's dgehrd has the same |subdiagonal|
        import scipy.linalg
        h2 = scipy.linalg.hessenberg(a)
        assert np.allclose(np.abs(np.diagonal(h2, -1)), np.abs(sub), rtol=1e-9, atol=1e-9)
    b = oracle.uniform(n, n, 22) * 2.0 - 1.0
    s = b @ b.T                                                                 # tridiagonal.rs:14
    tri, off = oracle.symmetric_tridiagonal(s)
    q = oracle.assemble_q(tri, off)
    rec = q @ _tri_matrix(np.diagonal(tri), np.abs(off)) @ q.T
    assert relative_eq(np.tril(s), np.tril(rec), 1e-7)
    assert np.allclose(np.linalg.eigvalsh(_tri_matrix(np.diagonal(tri), np.abs(off))), np.linalg.eigvalsh(s), atol=1e-9 * max(1, n))
    garbage = s.copy(); garbage[np.triu_indices(n, 1)] = np.nan                 # only the lower triangle is read
    tri2, off2 = oracle.symmetric_tridiagonal(garbage)
    assert np.array_equal(np.tril(tri2), np.tril(tri)) and np.array_equal(off2, off)
    s2 = s.copy(); s2[n // 2, :] = 0.0; s2[:, n // 2] = 0.0                     # tridiagonal.rs:22-31 (singular)
    tri, off = oracle.symmetric_tridiagonal(s2)
    q = oracle.assemble_q(tri, off)
    assert relative_eq(np.tril(s2), np.tril(q @ _tri_matrix(np.diagonal(tri), np.abs(off)) @ q.T), 1e-7)


@pytest.mark.parametrize("shape", RECT)
def test_oracle_bidiagonal_properties(oracle, shape):
    m, n = shape
    a = oracle.uniform(m, n, 23) * 200.0 - 100.0
    uv, d, e, upper = oracle.bidiagonal(a)
    assert upper == (m >= n)
    u, dm, vt = oracle.bidiagonal_u(uv, d, e), oracle.bidiagonal_d(d, e, upper), oracle.bidiagonal_v_t(uv, d, e)
    assert relative_eq(a, u @ dm @ vt, 1e-7 * np.abs(a).max())
    assert np.allclose(np.linalg.svd(dm, compute_uv=False), np.linalg.svd(a, compute_uv=False), rtol=1e-10, atol=1e-9)


def test_oracle_bidiagonal_identity_and_regressions(oracle):
    for m, n in [(10, 10), (10, 15), (15, 10)]:                                # bidiagonal.rs:62-78: assert_eq!
        a = np.asfortranarray(np.eye(m, n))
        uv, d, e, upper = oracle.bidiagonal(a)
        rec = oracle.bidiagonal_u(uv, d, e) @ oracle.bidiagonal_d(d, e, upper) @ oracle.bidiagonal_v_t(uv, d, e)
        assert np.array_equal(rec, a)
    s = float(np.float32(6.123234e-16))                                        # bidiagonal.rs:80-93 (issue 1313)
    a = np.array([[10.0, 0, 0, 0, -10, 0, 0, 0], [s, 10, 0, 10, s, 0, 0, 0], [20, -20, 0, 20, 20, 0, 0, 0]])
    a /= np.abs(a).max()
    uv, d, e, upper = oracle.bidiagonal(a)
    rec = oracle.bidiagonal_u(uv, d, e) @ oracle.bidiagonal_d(d, e, upper) @ oracle.bidiagonal_v_t(uv, d, e)
    assert np.allclose(a, rec, rtol=1e-6, atol=1e-6)
    s = float(np.float32(6.123234e-17))                                        # bidiagonal.rs:95-104
    a = np.array([[1.0, 0, -1], [s, 1, s]])
    uv, d, e, upper = oracle.bidiagonal(a)
    rec = oracle.bidiagonal_u(uv, d, e) @ oracle.bidiagonal_d(d, e, upper) @ oracle.bidiagonal_v_t(uv, d, e)
    assert np.allclose(a, rec, rtol=1e-6, atol=1e-6)


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_hessenberg_simple(nab):
    m = np.asfortranarray(np.array([[1.0, 0.0], [1.0, 3.0]]))
    p, h = nab.Hessenberg.new(m).unpack()
    assert relative_eq(m, p @ h @ p.T, 1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("n", SQ + [257, 600, 1100, 2100])
def test_hessenberg_vs_oracle(nab, oracle, n):
    a = oracle.uniform(n, n, 24) - 0.4
    got = nab.Hessenberg.new(a)
    if n <= 1100:
        hess_ref, sub_ref = oracle.hessenberg(a)
        tol = _storage_tol(n)
        assert np.abs(got.hess_internal() - hess_ref).max() <= tol
        assert np.abs(got.subdiag - sub_ref).max(initial=0.0) <= tol
    q, h = got.unpack()
    assert np.abs(np.tril(h, -2)).max(initial=0.0) == 0.0
    assert np.linalg.norm(q @ h @ q.T - a) <= 10 * n * EPS * np.linalg.norm(a)
    assert np.linalg.norm(q.T @ q - np.eye(n)) <= 10 * n * EPS


@pytest.mark.gpu
@pytest.mark.parametrize("n", SQ + [257, 600, 1100, 2100])
def test_symmetric_tridiagonal_vs_oracle(nab, oracle, n):
    b = oracle.uniform(n, n, 25) * 2.0 - 1.0
    s = np.asfortranarray((b + b.T) / 2.0)
    dirty = s.copy(order="F"); dirty[np.triu_indices(n, 1)] = np.nan           # the strict upper triangle is never read
    got = nab.SymmetricTridiagonal.new(dirty)
    if n <= 1100:
        tri_ref, off_ref = oracle.symmetric_tridiagonal(s)
        tol = _storage_tol(n)
        assert np.abs(np.tril(got.internal_tri()) - np.tril(tri_ref)).max() <= tol
        assert np.abs(got._off - off_ref).max(initial=0.0) <= tol
    assert np.all(np.isnan(got.internal_tri()[np.triu_indices(n, 1)]))         # ... and never written
    q, d, off = got.unpack()
    t = _tri_matrix(d, off)
    assert np.linalg.norm(q @ t @ q.T - s) <= 10 * n * EPS * np.linalg.norm(s)
    assert np.linalg.norm(q.T @ q - np.eye(n)) <= 10 * n * EPS
    assert np.allclose(np.linalg.eigvalsh(t), np.linalg.eigvalsh(s), atol=10 * n * EPS * np.linalg.norm(s))


@pytest.mark.gpu
def test_symmetric_tridiagonal_recompose_and_singular(nab, oracle):
    b = oracle.uniform(40, 40, 26) * 2.0 - 1.0
    s = np.asfortranarray(b @ b.T)
    s[20, :] = 0.0; s[:, 20] = 0.0                                             # tridiagonal.rs:22-31
    rec = nab.SymmetricTridiagonal.new(s).recompose()
    assert relative_eq(np.tril(s), np.tril(rec), 1e-7)
    z = nab.SymmetricTridiagonal.new(np.zeros((9, 9), order="F"))              # no reflection at all
    assert np.array_equal(z.off_diagonal(), np.zeros(8)) and np.array_equal(z.q(), np.eye(9))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", RECT + [(257, 300), (300, 257), (700, 512), (1100, 1100), (2100, 1030), (3000, 40), (40, 3000),
                                          (160000, 3), (3, 160000)])      # more row blocks than CTAs
def test_bidiagonal_vs_oracle(nab, oracle, shape):
    m, n = shape
    a = oracle.uniform(m, n, 27) - 0.4
    got = nab.Bidiagonal.new(a)
    assert got.is_upper_diagonal() == (m >= n)
    if max(m, n) <= 1100 or min(m, n) <= 3:
        uv_ref, d_ref, e_ref, _ = oracle.bidiagonal(a)
        tol = _storage_tol(max(m, n))
        assert np.abs(got.uv_internal() - uv_ref).max() <= tol
        assert np.abs(got._diag - d_ref).max() <= tol and np.abs(got._off - e_ref).max(initial=0.0) <= tol
    u, d, vt = got.unpack()
    mn = min(m, n)
    assert u.shape == (m, mn) and d.shape == (mn, mn) and vt.shape == (mn, n)
    assert np.linalg.norm(u @ d @ vt - a) <= 10 * max(m, n) * EPS * np.linalg.norm(a)
    assert np.linalg.norm(u.T @ u - np.eye(mn)) <= 10 * max(m, n) * EPS
    assert np.linalg.norm(vt @ vt.T - np.eye(mn)) <= 10 * max(m, n) * EPS
    assert np.allclose(np.linalg.svd(d, compute_uv=False), np.linalg.svd(a, compute_uv=False), atol=10 * max(m, n) * EPS * np.linalg.norm(a))


@pytest.mark.gpu
def test_bidiagonal_identity_and_regressions(nab):
    for m, n in [(10, 10), (10, 15), (15, 10)]:
        a = np.asfortranarray(np.eye(m, n))
        u, d, vt = nab.Bidiagonal.new(a).unpack()
        assert np.array_equal(u @ d @ vt, a)                                   # assert_eq! in the reference
    s = float(np.float32(6.123234e-16))
    a = np.array([[10.0, 0, 0, 0, -10, 0, 0, 0], [s, 10, 0, 10, s, 0, 0, 0], [20, -20, 0, 20, 20, 0, 0, 0]])
    a /= np.abs(a).max()
    u, d, vt = nab.Bidiagonal.new(a).unpack()
    assert np.allclose(a, u @ d @ vt, rtol=1e-6, atol=1e-6)
    s = float(np.float32(6.123234e-17))
    a = np.array([[1.0, 0, -1], [s, 1, s]])
    u, d, vt = nab.Bidiagonal.new(a).unpack()
    assert np.allclose(a, u @ d @ vt, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_two_sided_zero_columns_and_errors(nab, oracle):
    a = oracle.uniform(30, 30, 28) - 0.5
    a[3:, 2] = 0.0; a[:, 11] = 0.0                                             # a column that needs no reflection
    got = nab.Hessenberg.new(a); hess_ref, sub_ref = oracle.hessenberg(a)
    assert np.abs(got.hess_internal() - hess_ref).max() <= 1e-10 and np.abs(got.subdiag - sub_ref).max() <= 1e-10
    b = oracle.uniform(30, 12, 29) - 0.5; b[:, 0] = 0.0; b[7, :] = 0.0         # the first column needs no reflection
    gb = nab.Bidiagonal.new(b); uv_ref, d_ref, e_ref, _ = oracle.bidiagonal(b)
    assert np.abs(gb.uv_internal() - uv_ref).max() <= 1e-10 and np.abs(gb._diag - d_ref).max() <= 1e-10 and np.abs(gb._off - e_ref).max() <= 1e-10
    assert gb._diag[0] == 0.0 == d_ref[0]
    u, d, vt = gb.unpack()
    assert np.linalg.norm(u @ d @ vt - b) <= 1e-12
    for cls in (nab.Hessenberg, nab.SymmetricTridiagonal):
        with pytest.raises(ValueError):
            cls.new(np.zeros((3, 4)))
        with pytest.raises(ValueError):
            cls.new(np.zeros((0, 0)))
    with pytest.raises(ValueError):
        nab.Bidiagonal.new(np.zeros((0, 3)))


@pytest.mark.gpu
def test_two_sided_is_deterministic(nab, oracle):
    a = oracle.uniform(700, 700, 30) - 0.5
    h1, h2 = nab.Hessenberg.new(a), nab.Hessenberg.new(a)
    assert np.array_equal(h1.hess_internal(), h2.hess_internal()) and np.array_equal(h1.subdiag, h2.subdiag)
    b1, b2 = nab.Bidiagonal.new(a), nab.Bidiagonal.new(a)
    assert np.array_equal(b1.uv_internal(), b2.uv_internal())


@pytest.mark.gpu
def test_two_sided_4096_properties(nab, oracle):
    """Beyond the oracle's reach in test time: reconstruction and orthogonality at n = 4096 with the products on the GPU
    (north-star gates 10 n eps), plus the trace / Frobenius invariants of a similarity / orthogonal equivalence."""
    n = 4096
    a = oracle.uniform(n, n, 31) - 0.5
    h = nab.Hessenberg.new(a)
    q, hm = h.unpack()
    rec = nab.mul(nab.mul(q, hm), np.asfortranarray(q.T))
    assert np.linalg.norm(rec - a) <= 10 * n * EPS * np.linalg.norm(a)
    assert np.linalg.norm(nab.tr_mul(q, q) - np.eye(n)) <= 10 * n * EPS
    assert abs(np.trace(hm) - np.trace(a)) <= 10 * n * EPS * np.linalg.norm(a)
    s = np.asfortranarray((a + a.T) / 2.0)
    t = nab.SymmetricTridiagonal.new(s)
    rec = t.recompose()
    assert np.linalg.norm(np.tril(rec) - np.tril(s)) <= 10 * n * EPS * np.linalg.norm(s)
    b = nab.Bidiagonal.new(a)
    u, d, vt = b.unpack()
    rec = nab.mul(nab.mul(u, d), vt)
    assert np.linalg.norm(rec - a) <= 10 * n * EPS * np.linalg.norm(a)
    assert abs(np.linalg.norm(d) - np.linalg.norm(a)) <= 10 * n * EPS * np.linalg.norm(a)


@pytest.mark.gpu
def test_wire_round_trip_of_gpu_factors(nab, oracle):
    """A factor object computed on the GPU survives nalgebra's serde wire format bit for bit and still solves."""
    a = oracle.uniform(300, 300, 32) - 0.5; b = oracle.uniform(300, 2, 33)
    lu = nab.LU.new(a)
    back = nab.wire.loads(nab.wire.dumps(lu), "LU")
    assert np.array_equal(back.lu, lu.lu) and np.array_equal(back.p().ipiv, lu.p().ipiv)
    assert np.array_equal(back.solve(b), lu.solve(b))
    hs = nab.Hessenberg.new(a)
    back = nab.wire.loads(nab.wire.dumps(hs), "Hessenberg")
    assert np.array_equal(back.q(), hs.q()) and np.array_equal(back.h(), hs.h())


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 3, 7, 64, 130, 600, 1100, 2100])
def test_fused_and_two_pass_kernels_agree(nab, oracle, n):
    """Hessenberg / SymmetricTridiagonal have two kernels (one fused pass per step from n = 6144 / 16384, two passes
    below): both against the oracle at oracle sizes, and against each other."""
    from nalgebra_b200 import _capi
    L = _capi.lib()
    a = oracle.uniform(n, n, 34) - 0.4
    sy = np.asfortranarray((a + a.T) / 2.0)
    res = {}
    try:
        for mode in (0, 1):
            _capi.check(L.na_set_tuning(b"ts_fused", mode))
            res[mode] = (nab.Hessenberg.new(a), nab.SymmetricTridiagonal.new(sy))
    finally:
        _capi.check(L.na_set_tuning(b"ts_fused", -1))
    tol = _storage_tol(n)
    assert np.abs(res[0][0].hess_internal() - res[1][0].hess_internal()).max() <= tol
    assert np.abs(np.tril(res[0][1].internal_tri()) - np.tril(res[1][1].internal_tri())).max() <= tol
    if n <= 1100:
        hess_ref, sub_ref = oracle.hessenberg(a)
        tri_ref, off_ref = oracle.symmetric_tridiagonal(sy)
        assert np.abs(res[1][0].hess_internal() - hess_ref).max() <= tol and np.abs(res[1][0].subdiag - sub_ref).max() <= tol
        assert np.abs(np.tril(res[1][1].internal_tri()) - np.tril(tri_ref)).max() <= tol and np.abs(res[1][1]._off - off_ref).max() <= tol
    q, h = res[1][0].unpack()
    assert np.linalg.norm(q @ h @ q.T - a) <= 10 * n * EPS * np.linalg.norm(a)
    rec = res[1][1].recompose()
    assert np.linalg.norm(np.tril(rec) - np.tril(sy)) <= 10 * n * EPS * np.linalg.norm(sy)
