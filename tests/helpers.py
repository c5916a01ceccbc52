"""Shared helpers: KAT loading, the reference's test generators restated, tolerances."""
import json
import os

import numpy as np

EPS = np.finfo(np.float64).eps
HERE = os.path.dirname(os.path.abspath(__file__))


def load_kats():
    with open(os.path.join(HERE, "golden", "nalgebra_kats.json")) as f:
        return json.load(f)


def mat(rows):
    """Row-major nested list (as written in the Rust source) -> column-major float64 matrix."""
    return np.asfortranarray(np.array([[float("nan") if v == "nan" else v for v in r] for r in rows], dtype=np.float64))


def random_orthogonal(n, rng):
    """RandomOrthogonal: product of random Givens rotations (src/debug/random_orthogonal.rs:32-42)."""
    q = np.eye(n)
    if n == 1:
        return q * (1.0 if rng.random() < 0.5 else -1.0)
    for i in range(n - 1):
        c, s = rng.random(), rng.random()
        nrm = np.hypot(c, s)
        c, s = c / nrm, s / nrm
        g = np.eye(n)
        g[i, i] = c; g[i, i + 1] = -s; g[i + 1, i] = s; g[i + 1, i + 1] = c
        q = g @ q
    return q


def random_sdp(n, rng):
    """RandomSDP: Q * diag(1 + |r|) * Q^T, eigenvalues in [1, 2) (src/debug/random_sdp.rs:34-45)."""
    q = random_orthogonal(n, rng)
    d = 1.0 + np.abs(rng.random(n))
    return np.asfortranarray((q * d) @ q.T)


def bench_spd(n, rng):
    """The reference benches' SPD recipe: M*M^T + sqrt(eps)*|M|_F^2 * I (benches/linalg/cholesky.rs:3-11)."""
    m = rng.random((n, n))
    return np.asfortranarray(m @ m.T + np.sqrt(EPS) * np.linalg.norm(m) ** 2 * np.eye(n))


def gemm_tol(a, b, k):
    """north_star: max element error <= 4*k*eps*|A||B| (Frobenius norms)."""
    return 4 * max(k, 1) * EPS * np.linalg.norm(a) * np.linalg.norm(b) + 1e-300


def relative_eq(a, b, epsilon):
    """approx::relative_eq!(a, b, epsilon = e): elementwise |a-b| <= e or <= max_relative(=eps) * max(|a|,|b|)."""
    a = np.asarray(a); b = np.asarray(b)
    d = np.abs(a - b)
    return bool(np.all((d <= epsilon) | (d <= EPS * np.maximum(np.abs(a), np.abs(b)))))


def first_pivot_divergence(a, lu_packed, swaps_a, swaps_b):
    """Tie-aware comparison of two PermutationSequences of the same matrix.  None when they are equal; otherwise the
    first step at which they differ, the two pivot rows, their candidate values in that column (recomputed from the
    common prefix of the factorization: s = (P A)[i:, i] - L[i:, :i] U[:i, i]) and the relative gap between them --
    a gap of a few ulps means a legitimate near-tie resolved differently by a different summation order, anything
    larger is a bug.  `lu_packed` must belong to `swaps_a`."""
    sa = [tuple(int(v) for v in p) for p in np.asarray(swaps_a).reshape(-1, 2)]
    sb = [tuple(int(v) for v in p) for p in np.asarray(swaps_b).reshape(-1, 2)]
    if sa == sb:
        return None
    n = a.shape[0]
    da, db = dict(sa), dict(sb)                      # step -> pivot row (missing: no swap, pivot = step)
    step = min(i for i in range(min(a.shape)) if da.get(i, i) != db.get(i, i))
    perm = np.arange(n)
    for i, j in sa:
        if i >= step:
            break
        perm[[i, j]] = perm[[j, i]]
    pa = a[perm]
    l = np.tril(lu_packed[:, :step], -1)[step:, :]   # rows >= step of L's first `step` columns (A's permutation beyond `step` does not matter
    # for the candidates' VALUES, only for their positions: use positions of sequence a up to `step`, which both share)
    u = np.triu(lu_packed[:step, :])[:, step] if step else np.zeros(0)
    # rows >= step of lu_packed are permuted by a's later swaps; undo them to address rows by their position at `step`
    later = [(i, j) for i, j in sa if i >= step]
    rowpos = np.arange(n)
    for i, j in later:
        rowpos[[i, j]] = rowpos[[j, i]]
    lfull = np.empty((n - step, step))
    lfull[rowpos[step:] - step] = l                  # row r of the final factor sat at position rowpos[r] at `step`
    s = pa[step:, step] - lfull @ u
    ra, rb = da.get(step, step), db.get(step, step)
    va, vb = s[ra - step], s[rb - step]
    return {"step": step, "rows": (ra, rb), "values": (float(va), float(vb)),
            "relative_gap": float(abs(abs(va) - abs(vb)) / max(abs(va), abs(vb), 1e-300))}
