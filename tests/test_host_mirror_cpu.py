"""Host-side logic of the two-sided reductions' mirrors (nalgebra_b200/linalg.py: `q()` / `u()` / `v_t()` through one
`na_qr_q_f64` call on the shifted or transposed storage, `h()`, `d()`, `recompose()`) on the CPU box: the C-ABI entry
points they call are replaced by the oracle functions of the same signature (test double -- the product itself never
sees the oracle), and the results must equal the oracle's direct restatements of `householder::assemble_q`,
`Bidiagonal::u` and `Bidiagonal::v_t` (src/linalg/householder.rs:132-152, bidiagonal.rs:205-283)."""
import numpy as np
import pytest


class _OracleBackedLib:
    """Same argument lists as include/nalgebra_b200.h for the calls the mirrors make."""
    _VOID = {"na_hessenberg_f64": "na_oracle_hessenberg_f64", "na_symmetric_tridiagonal_f64": "na_oracle_symmetric_tridiagonal_f64",
             "na_bidiagonal_f64": "na_oracle_bidiagonal_f64", "na_qr_q_f64": "na_oracle_qr_q_f64", "na_dgemm": "na_oracle_gemm_f64",
             "na_lu_f64": "na_oracle_lu_f64", "na_qr_f64": "na_oracle_qr_f64", "na_qr_q_tr_mul_f64": "na_oracle_qr_q_tr_mul_f64",
             "na_full_piv_lu_f64": "na_oracle_full_piv_lu_f64", "na_col_piv_qr_f64": "na_oracle_col_piv_qr_f64",
             "na_cholesky_solve_f64": "na_oracle_cholesky_solve_f64"}
    _BOOL = {"na_lu_solve_f64": "na_oracle_lu_solve_f64", "na_qr_solve_f64": "na_oracle_qr_solve_f64"}   # oracle: 1 = solved, 0 = singular

    def __init__(self, O):
        self._o = O.lib()

    def na_cholesky_f64(self, *args):                      # 0 / 1 on both sides (NA_OK / NA_NOT_PD)
        return self._o.na_oracle_cholesky_f64(*args)

    def na_tri_solve_f64(self, lower, trans, unit, n, t, ldt, b, ldb, nrhs):
        import ctypes as C
        import scipy.linalg
        tm = np.ctypeslib.as_array((C.c_double * (ldt * n)).from_address(t)).reshape((ldt, n), order="F")[:n]
        bm = np.ctypeslib.as_array((C.c_double * (ldb * nrhs)).from_address(b)).reshape((ldb, nrhs), order="F")
        if not unit and np.any(np.diagonal(tm) == 0.0):
            return 2                                       # NA_SINGULAR
        bm[:n] = scipy.linalg.solve_triangular(tm, bm[:n], lower=bool(lower), trans=int(trans), unit_diagonal=bool(unit))
        return 0

    def __getattr__(self, name):
        if name in self._VOID:
            fn = getattr(self._o, self._VOID[name])
            return lambda *a: (fn(*a), 0)[1]
        if name in self._BOOL:
            fn = getattr(self._o, self._BOOL[name])
            return lambda *a: 0 if fn(*a) else 2
        raise AttributeError(name)


@pytest.fixture()
def nab_on_oracle(monkeypatch, oracle):
    import nalgebra_b200
    from nalgebra_b200 import _capi
    fake = _OracleBackedLib(oracle)
    monkeypatch.setattr(_capi, "lib", lambda: fake)
    return nalgebra_b200


@pytest.mark.parametrize("n", [1, 2, 3, 6, 17, 40])
def test_hessenberg_and_tridiagonal_mirrors(nab_on_oracle, oracle, n):
    nab = nab_on_oracle
    a = oracle.uniform(n, n, 41) - 0.4
    h = nab.Hessenberg.new(a)
    hess_ref, sub_ref = oracle.hessenberg(a)
    assert np.array_equal(h.hess_internal(), hess_ref) and np.array_equal(h.subdiag, sub_ref)
    assert np.array_equal(h.q(), oracle.assemble_q(hess_ref, sub_ref))          # 1 (+) QR::q of the storage one row down
    assert np.array_equal(h.h(), oracle.hessenberg_h(hess_ref, sub_ref))
    q, hm = h.unpack()
    assert np.abs(q @ hm @ q.T - a).max() <= 1e-12
    s = np.asfortranarray((a + a.T) / 2)
    t = nab.SymmetricTridiagonal.new(s)
    tri_ref, off_ref = oracle.symmetric_tridiagonal(s)
    assert np.array_equal(t.q(), oracle.assemble_q(tri_ref, off_ref))
    assert np.array_equal(t.off_diagonal(), np.abs(off_ref)) and np.array_equal(t.diagonal(), np.diagonal(tri_ref))
    assert np.abs(np.tril(t.recompose()) - np.tril(s)).max() <= 1e-12


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (5, 3), (3, 5), (7, 1), (1, 7), (10, 15), (15, 10), (12, 12)])
def test_bidiagonal_mirror(nab_on_oracle, oracle, shape):
    nab = nab_on_oracle
    m, n = shape
    a = oracle.uniform(m, n, 42) - 0.4
    b = nab.Bidiagonal.new(a)
    uv, d, e, upper = oracle.bidiagonal(a)
    assert b.is_upper_diagonal() == upper and np.array_equal(b.uv_internal(), uv)
    assert np.array_equal(b.u(), oracle.bidiagonal_u(uv, d, e))                 # shift 0: QR::q; shift 1: 1 (+) q of the shifted storage
    # the same on the transposed storage; the reference sweeps rows (gemv + ger), the mirror columns: equal to rounding
    assert np.abs(b.v_t() - oracle.bidiagonal_v_t(uv, d, e)).max() <= 1e-14
    assert np.array_equal(b.d(), oracle.bidiagonal_d(d, e, upper))
    u, dm, vt = b.unpack()
    assert np.abs(u @ dm @ vt - a).max() <= 1e-12


def test_factorization_mirrors_host_logic(nab_on_oracle, oracle):
    """Accessors, determinants, solve / inverse wrappers and the order in which the permutation sequences are applied
    (cholesky.rs:82-185, lu.rs:132-331, qr.rs:81-294, full_piv_lu.rs:94-270, col_piv_qr.rs:97-337)."""
    nab = nab_on_oracle
    n = 12
    a = oracle.uniform(n, n, 43) - 0.5
    b = oracle.uniform(n, 3, 44)
    det = np.linalg.det(a)
    x_ref = np.linalg.solve(a, b)
    for f in (nab.LU.new(a), nab.FullPivLU.new(a), nab.QR.new(a), nab.ColPivQR.new(a)):
        x = f.solve(b)
        assert x is not None and np.abs(x - x_ref).max() <= 1e-10, type(f).__name__
        assert np.abs(f.try_inverse() @ a - np.eye(n)).max() <= 1e-10
        assert f.is_invertible()
        if isinstance(f, nab.ColPivQR):
            # col_piv_qr.rs:324-337 multiplies the SIGNED diag entries and p.determinant(): the magnitude is the determinant's,
            # the sign follows the reference's formula (it ignores the reflections' own determinants), and so does the mirror
            assert abs(abs(f.determinant()) / abs(det) - 1.0) <= 1e-10
            assert f.determinant() == np.prod(f.diag) * f.p().determinant()
        elif not isinstance(f, nab.QR):
            assert abs(f.determinant() / det - 1.0) <= 1e-10, type(f).__name__
        xv = f.solve(b[:, 0])                                                    # a DVector right-hand side
        assert xv.shape == (n,) and np.abs(xv - x_ref[:, 0]).max() <= 1e-10
    lu = nab.LU.new(a)
    pm, l, u = lu.unpack()
    rec = l @ u
    pm.inv_permute_rows(rec)
    assert np.abs(rec - a).max() <= 1e-13
    fp = nab.FullPivLU.new(a)
    p, l, u, q = fp.unpack()
    rec = l @ u
    p.inv_permute_rows(rec); q.inv_permute_columns(rec)
    assert np.abs(rec - a).max() <= 1e-13
    cp = nab.ColPivQR.new(a)
    qm, r, pc = cp.unpack()
    rec = qm @ r
    pc.inv_permute_columns(rec)
    assert np.abs(rec - a).max() <= 1e-13
    spd = oracle.spd_wellcond(n, 45)
    ch = nab.Cholesky.new(spd)
    assert np.abs(ch.l() @ ch.l().T - spd).max() <= 1e-12 and np.abs(ch.solve(b) - np.linalg.solve(spd, b)).max() <= 1e-12
    assert abs(ch.determinant() / np.linalg.det(spd) - 1.0) <= 1e-10 and abs(ch.ln_determinant() - np.log(np.linalg.det(spd))) <= 1e-9
    assert np.abs(ch.inverse() @ spd - np.eye(n)).max() <= 1e-12
    assert nab.Cholesky.new(-spd) is None
    sing = a.copy(); sing[:, 3] = 0.0                                          # exactly singular: only an exact zero pivot makes the reference return None
    assert nab.LU.new(sing).solve(b) is None and not nab.FullPivLU.new(sing).is_invertible() and nab.FullPivLU.new(sing).determinant() == 0.0


def test_matrix_level_entry_points(nab_on_oracle, oracle):
    """decomposition.rs / determinant.rs:20-58 / inverse.rs:16-150: closed forms up to dimension 3 (4 for the inverse), LU
    beyond; the 5 x 5 table of tests/linalg/inverse.rs:62-81 (relative 1e-4 there)."""
    nab = nab_on_oracle
    for n in range(0, 9):
        a = oracle.uniform(n, n, 46 + n) - 0.5
        assert abs(nab.determinant(a) - (np.linalg.det(a) if n else 1.0)) <= 1e-12
        inv = nab.try_inverse(a)
        assert inv is not None and inv.shape == (n, n) and (n == 0 or np.abs(inv @ a - np.eye(n)).max() <= 1e-9)
    for n in (1, 2, 3, 4, 6):
        z = np.zeros((n, n)); z[:, : n - 1] = oracle.uniform(n, n - 1, 3) if n > 1 else z[:, :0]
        assert nab.try_inverse(z) is None and nab.determinant(z) == 0.0
    with pytest.raises(ValueError):
        nab.determinant(np.zeros((2, 3)))
    a = oracle.uniform(7, 7, 60) - 0.5
    assert isinstance(nab.lu(a), nab.LU) and isinstance(nab.qr(a), nab.QR) and isinstance(nab.hessenberg(a), nab.Hessenberg)
    assert isinstance(nab.bidiagonalize(a), nab.Bidiagonal) and isinstance(nab.full_piv_lu(a), nab.FullPivLU)
    assert isinstance(nab.col_piv_qr(a), nab.ColPivQR) and isinstance(nab.symmetric_tridiagonalize(a + a.T), nab.SymmetricTridiagonal)
    assert nab.cholesky(a @ a.T + 7 * np.eye(7)) is not None and nab.cholesky(-np.eye(7)) is None
    h = nab.Hessenberg.new_with_workspace(a, np.zeros(7))
    assert np.array_equal(h.hess_internal(), nab.hessenberg(a).hess_internal())
    with pytest.raises(ValueError):
        nab.Hessenberg.new_with_workspace(a, np.zeros(6))
