"""The serde wire format of nalgebra's dynamic matrices and factorization objects, so that factor
objects computed on the GPU can be handed to a nalgebra process and back.

Reference: ``VecStorage`` serializes as the tuple ``(data, nrows, ncols)`` with ``data`` in column-major order
(/root/reference/src/base/vec_storage.rs:74-114; deserialization rejects ``nrows * ncols != len(data)``), ``Matrix`` as
its storage (src/base/matrix.rs:240-253), ``Dyn`` as its integer and ``Const<N>`` as the unit value
(src/base/dimension.rs:41-48, 238-245) -- so with serde_json a ``DMatrix`` is ``[[...], nrows, ncols]`` and a ``DVector``
``[[...], nrows, null]``.  The factorization structs derive ``Serialize`` field by field
(src/linalg/{cholesky,lu,qr,full_piv_lu,col_piv_qr,hessenberg,symmetric_tridiagonal,bidiagonal}.rs,
permutation_sequence.rs:28-34: ``{len, ipiv}`` with ``ipiv`` a vector of ``(usize, usize)`` pairs of full capacity).

Golden vector held by the reference: tests/core/serde.rs:30-34, ``[[1.0, 2.0, 3.0, 4.0, 5.0, 6.0],2,3]``.
"""
from __future__ import annotations

import json

import numpy as np

from . import linalg as _l


def dmatrix_to_wire(m) -> list:
    m = np.asarray(m, dtype=np.float64)
    if m.ndim != 2:
        raise ValueError("expected a matrix")
    return [m.reshape(-1, order="F").tolist(), int(m.shape[0]), int(m.shape[1])]


def dmatrix_from_wire(w) -> np.ndarray:
    data, nrows, ncols = w
    if nrows * ncols != len(data):                                     # vec_storage.rs:103-109
        raise ValueError(f"Expected {nrows * ncols} components, found {len(data)}")
    return np.array(data, dtype=np.float64).reshape((nrows, ncols), order="F")


def dvector_to_wire(v) -> list:
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    return [v.tolist(), int(v.shape[0]), None]


def dvector_from_wire(w) -> np.ndarray:
    data, nrows, ncols = w
    if ncols is not None:
        raise ValueError("a DVector carries the unit value as its column dimension")
    if nrows != len(data):
        raise ValueError(f"Expected {nrows} components, found {len(data)}")
    return np.array(data, dtype=np.float64)


def permutation_to_wire(p: "_l.PermutationSequence") -> dict:
    pairs = [[int(a), int(b)] for a, b in p.ipiv] + [[0, 0]] * (p.capacity - len(p))       # capacity entries, (0, 0) beyond len
    return {"len": len(p), "ipiv": [pairs, int(p.capacity), None]}


def permutation_from_wire(w) -> "_l.PermutationSequence":
    pairs, cap, ncols = w["ipiv"]
    if ncols is not None or cap != len(pairs) or w["len"] > cap:
        raise ValueError("malformed PermutationSequence")
    return _l.PermutationSequence(np.array(pairs[: w["len"]], dtype=np.uint64).reshape(-1, 2), cap)


# field name -> (attribute on the host mirror, kind)
_LAYOUT = {
    "Cholesky": (_l.Cholesky, [("chol", "chol", "m")]),
    "LU": (_l.LU, [("lu", "lu", "m"), ("p", "_p", "p")]),
    "QR": (_l.QR, [("qr", "qr", "m"), ("diag", "diag", "v")]),
    "FullPivLU": (_l.FullPivLU, [("lu", "lu", "m"), ("p", "_p", "p"), ("q", "_q", "p")]),
    "ColPivQR": (_l.ColPivQR, [("col_piv_qr", "col_piv_qr", "m"), ("p", "_p", "p"), ("diag", "diag", "v")]),
    "Hessenberg": (_l.Hessenberg, [("hess", "hess", "m"), ("subdiag", "subdiag", "v")]),
    "SymmetricTridiagonal": (_l.SymmetricTridiagonal, [("tri", "tri", "m"), ("off_diagonal", "_off", "v")]),
    "Bidiagonal": (_l.Bidiagonal, [("uv", "uv", "m"), ("diagonal", "_diag", "v"), ("off_diagonal", "_off", "v"),
                                   ("upper_diagonal", "upper_diagonal", "b")]),
}
_TO = {"m": dmatrix_to_wire, "v": dvector_to_wire, "p": permutation_to_wire, "b": bool}
_FROM = {"m": lambda w: np.asfortranarray(dmatrix_from_wire(w)), "v": dvector_from_wire, "p": permutation_from_wire, "b": bool}


def to_wire(obj) -> dict:
    """A factorization object -> the value serde would produce for nalgebra's struct of the same name."""
    name = type(obj).__name__
    if name not in _LAYOUT:
        raise TypeError(f"no wire format for {name}")
    return {field: _TO[kind](getattr(obj, attr)) for field, attr, kind in _LAYOUT[name][1]}


def from_wire(name: str, w: dict):
    cls, fields = _LAYOUT[name]
    obj = cls.__new__(cls)
    for field, attr, kind in fields:
        setattr(obj, attr, _FROM[kind](w[field]))
    return obj


def dumps(obj) -> str:
    """serde_json::to_string of a DMatrix (2-D array), DVector (1-D array) or factorization object."""
    if isinstance(obj, np.ndarray):
        return json.dumps(dmatrix_to_wire(obj) if obj.ndim == 2 else dvector_to_wire(obj), separators=(",", ":"))
    return json.dumps(to_wire(obj), separators=(",", ":"))


def loads(text: str, name: str = "DMatrix"):
    w = json.loads(text)
    if name == "DMatrix":
        return dmatrix_from_wire(w)
    if name == "DVector":
        return dvector_from_wire(w)
    return from_wire(name, w)
