"""The serde wire format against the golden strings the reference's own tests hold
(/root/reference/tests/core/serde.rs:30-34, 44-48) and round trips of every factorization object; host logic only."""
import json

import numpy as np
import pytest

from nalgebra_b200 import linalg as L
from nalgebra_b200 import wire


def test_dmatrix_golden_string():
    m = wire.loads("[[1.0, 2.0, 3.0, 4.0, 5.0, 6.0],2,3]")                      # serde.rs:30-34
    assert np.array_equal(m, np.array([[1.0, 3.0, 5.0], [2.0, 4.0, 6.0]]))    # from_column_slice(2, 3, ..)
    assert json.loads(wire.dumps(m)) == [[1.0, 2.0, 3.0, 4.0, 5.0, 6.0], 2, 3]
    with pytest.raises(ValueError, match="Expected 6 components, found 5"):     # serde.rs:44-48 (must fail)
        wire.loads("[[1.0, 2.0, 3.0, 4.0, 5.0],2,3]")


def test_dvector_and_permutation():
    v = np.array([1.5, -2.0, 0.25])
    assert json.loads(wire.dumps(v)) == [[1.5, -2.0, 0.25], 3, None]           # Const<1> is the unit value
    assert np.array_equal(wire.loads(wire.dumps(v), "DVector"), v)
    p = L.PermutationSequence(np.array([[0, 2], [1, 3]], dtype=np.uint64), 4)
    w = wire.permutation_to_wire(p)
    assert w == {"len": 2, "ipiv": [[[0, 2], [1, 3], [0, 0], [0, 0]], 4, None]}
    q = wire.permutation_from_wire(json.loads(json.dumps(w)))
    assert np.array_equal(q.ipiv, p.ipiv) and q.capacity == 4


def test_factorization_round_trips():
    rng = np.random.default_rng(0)
    a = np.asfortranarray(rng.random((4, 3)))
    p = L.PermutationSequence(np.array([[0, 2]], dtype=np.uint64), 3)
    objs = [L.Cholesky(np.asfortranarray(rng.random((3, 3)))), L.LU(a.copy(order="F"), p), L.QR(a.copy(order="F"), rng.random(3)),
            L.FullPivLU(a.copy(order="F"), p, L.PermutationSequence.identity(3)), L.ColPivQR(a.copy(order="F"), p, rng.random(3)),
            L.Hessenberg(np.asfortranarray(rng.random((3, 3))), rng.random(2)),
            L.SymmetricTridiagonal(np.asfortranarray(rng.random((3, 3))), rng.random(2)),
            L.Bidiagonal(a.copy(order="F"), rng.random(3), rng.random(2), True)]
    fields = {"Cholesky": ["chol"], "LU": ["lu", "p"], "QR": ["qr", "diag"], "FullPivLU": ["lu", "p", "q"],
              "ColPivQR": ["col_piv_qr", "p", "diag"], "Hessenberg": ["hess", "subdiag"],
              "SymmetricTridiagonal": ["tri", "off_diagonal"], "Bidiagonal": ["uv", "diagonal", "off_diagonal", "upper_diagonal"]}
    for obj in objs:
        name = type(obj).__name__
        text = wire.dumps(obj)
        assert list(json.loads(text).keys()) == fields[name]                   # the reference's field names and order
        back = wire.loads(text, name)
        for k, v in vars(obj).items():
            w = getattr(back, k)
            if isinstance(v, np.ndarray):
                assert np.array_equal(v, w) and (v.ndim != 2 or w.flags.f_contiguous)
            elif isinstance(v, L.PermutationSequence):
                assert np.array_equal(v.ipiv, w.ipiv) and v.capacity == w.capacity
            else:
                assert v == w
    lu = wire.loads(wire.dumps(objs[1]), "LU")
    assert json.loads(wire.dumps(lu))["lu"][1:] == [4, 3]
