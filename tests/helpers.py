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
