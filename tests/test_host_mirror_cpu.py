"""Host-side logic of the two-sided reductions' mirrors (nalgebra_b200/linalg.py: `q()` / `u()` / `v_t()` through one
`na_qr_q_f64` call on the shifted or transposed storage, `h()`, `d()`, `recompose()`) on the CPU box: the C-ABI entry
points they call are replaced by the oracle functions of the same signature (test double -- the product itself never
sees the oracle), and the results must equal the oracle's direct restatements of `householder::assemble_q`,
`Bidiagonal::u` and `Bidiagonal::v_t` (src/linalg/householder.rs:132-152, bidiagonal.rs:205-283)."""
import numpy as np
import pytest


class _OracleBackedLib:
    """Same argument lists as include/nalgebra_b200.h for the five calls the mirrors make."""

    def __init__(self, O):
        self._o = O.lib()

    def _void(self, fn):
        def call(*args):
            fn(*args)
            return 0
        return call

    def __getattr__(self, name):
        table = {"na_hessenberg_f64": "na_oracle_hessenberg_f64", "na_symmetric_tridiagonal_f64": "na_oracle_symmetric_tridiagonal_f64",
                 "na_bidiagonal_f64": "na_oracle_bidiagonal_f64", "na_qr_q_f64": "na_oracle_qr_q_f64", "na_dgemm": "na_oracle_gemm_f64"}
        if name not in table:
            raise AttributeError(name)
        return self._void(getattr(self._o, table[name]))


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
