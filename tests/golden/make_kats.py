"""Writes tests/golden/nalgebra_kats.json: the known-answer tests the REFERENCE's own test-suite and
doc-tests hold for the hot path, transcribed by hand from /root/reference (the reference is Rust and
cannot be executed in this image, so these are its literal vectors, not outputs of a run).

Each entry cites the reference file:line it was copied from.  Matrices are row-major nested lists
exactly as written in the Rust source (`MatrixRxC::new(...)` takes row-major arguments).

    python tests/golden/make_kats.py
"""
import json
import os

NAN = float("nan")
kats = {
    "gemm_doc": {
        "source": "src/base/blas.rs:713-727 (doc-test of Matrix::gemm)",
        "mat1": [[1, 0, 0, 0], [0, 1, 0, 0]],
        "mat2": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]],
        "mat3": [[0.1, 0.2, 0.3, 0.4], [0.5, 0.6, 0.7, 0.8], [0.9, 1.0, 1.1, 1.2]],
        "alpha": 10.0, "beta": 5.0,
        "check": "assert_relative_eq!(mat1, mat2 * mat3 * 10.0 + mat1 * 5.0)",
    },
    "gemm_tr_doc": {
        "source": "src/base/blas.rs:753-768 (doc-test of Matrix::gemm_tr)",
        "mat1": [[1, 0, 0, 0], [0, 1, 0, 0]],
        "mat2": [[1.0, 4.0], [2.0, 5.0], [3.0, 6.0]],
        "mat3": [[0.1, 0.2, 0.3, 0.4], [0.5, 0.6, 0.7, 0.8], [0.9, 1.0, 1.1, 1.2]],
        "alpha": 10.0, "beta": 5.0,
        "check": "assert_eq!(mat1, mat2.transpose() * mat3 * 10.0 + mat1 * 5.0)",
    },
    "axcpy_doc": {
        "source": "src/base/blas.rs:283-288",
        "y": [1.0, 2.0, 3.0], "x": [0.1, 0.2, 0.3], "a": 5.0, "c": 2.0, "b": 5.0, "expected": [6.0, 12.0, 18.0],
    },
    "axpy_doc": {
        "source": "src/base/blas.rs:306-311",
        "y": [1.0, 2.0, 3.0], "x": [0.1, 0.2, 0.3], "a": 10.0, "b": 5.0, "expected": [6.0, 12.0, 18.0],
    },
    "dot_doc": {
        "source": "src/base/blas.rs:176-180",
        "x": [1.0, 2.0, 3.0], "y": [0.1, 0.2, 0.3], "expected": 1.4,
    },
    "simple_mul": {
        "source": "tests/core/matrix.rs:421-435",
        "a": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]],
        "b": [[10.0, 20.0, 30.0, 40.0], [50.0, 60.0, 70.0, 80.0], [90.0, 100.0, 110.0, 120.0]],
        "expected": [[380.0, 440.0, 500.0, 560.0], [830.0, 980.0, 1130.0, 1280.0]],
    },
    "empty_matrix_mul_matrix": {
        "source": "tests/core/empty.rs:11-21", "shapes": [[3, 0, 4], [13, 0, 14]], "expected": "zeros",
    },
    "empty_matrix_gemm": {
        "source": "tests/core/empty.rs:37-50", "shapes": [[3, 0, 4], [13, 0, 14]], "c_init": 1.0, "alpha": 1.0, "beta": 0.5,
        "expected_fill": 0.5,
    },
    "empty_matrix_gemm_tr": {
        "source": "tests/core/empty.rs:53-60", "shape_a": [0, 3], "shape_b": [0, 4], "c_init": 1.0, "alpha": 1.0, "beta": 0.5,
        "expected_fill": 0.5,
    },
    "gemm_noncommutative": {
        "source": "tests/core/blas.rs:4-22",
        "note": "Quaternion scalars: exercises the generic gemv fallback's multiplication ORDER (alpha*a*b) and the "
                "beta path.  Not representable in f64; the f64 analogue kept here is the order-sensitive identity "
                "res.gemm(1, m1, m2, 0) == I and res.gemm(k, m1, m2, -k) == 0 for m2 = inverse(m1), on real 2x2 "
                "matrices with exactly representable entries.",
        "m1": [[2.0, 0.0], [4.0, 0.5]], "m2": [[0.5, 0.0], [-4.0, 2.0]], "k": 3.0,
    },
    "cholesky_with_substitute": {
        "source": "tests/linalg/cholesky.rs:3-12",
        "m": [[1.0, "nan"], [1.0, 1e-32]], "substitute": 1e-8,
        "expected": {"new": None, "new_with_substitute": "some"},
    },
    "lu_simple": {
        "source": "tests/linalg/lu.rs:3-20",
        "m": [[2.0, -1.0, 0.0], [-1.0, 2.0, -1.0], [0.0, -1.0, 2.0]], "determinant": 4.0, "epsilon": 1.0e-7,
    },
    "lu_simple_with_pivot": {
        "source": "tests/linalg/lu.rs:22-39",
        "m": [[0.0, -1.0, 2.0], [-1.0, 2.0, -1.0], [2.0, -1.0, 0.0]], "determinant": -4.0, "epsilon": 1.0e-7,
    },
    "matrix5_try_inverse": {
        "source": "tests/linalg/inverse.rs:62-81 (Matrix::try_inverse -> lu::try_invert_to)",
        "a": [[-2.0, 0.0, 2.0, 5.0, -5.0], [-6.0, 4.0, 4.0, 13.0, -15.0], [4.0, 16.0, -14.0, -19.0, 12.0],
              [12.0, 12.0, -22.0, -35.0, 34.0], [-8.0, 4.0, 12.0, 27.0, -31.0]],
        "expected_inverse": [[3.9333e+00, -1.5667e+00, 2.6667e-01, 6.6667e-02, 3.0000e-01],
                             [-1.2033e+01, 3.9667e+00, -1.1167e+00, 2.8333e-01, -1.0000e-01],
                             [-1.8233e+01, 5.7667e+00, -1.5667e+00, 2.3333e-01, -2.0000e-01],
                             [-4.3333e+00, 1.6667e+00, -6.6667e-01, 3.3333e-01, -4.6950e-19],
                             [-1.3400e+01, 4.6000e+00, -1.4000e+00, 4.0000e-01, -2.0000e-01]],
        "max_relative": 1e-4,
    },
    "proptest_params": {
        "source": "tests/proptest/mod.rs:19-20; tests/linalg/{cholesky,lu,qr}.rs",
        "matrix_dim": [1, 20], "f64_range": [-100.0, 100.0],
        "tolerances": {"cholesky_recompose": 1e-7, "cholesky_solve": 1e-7, "lu_recompose": 1e-7, "lu_solve": 1e-6,
                       "lu_inverse_identity": 1e-5, "qr_recompose": 1e-7, "qr_orthogonal": 1e-7, "qr_solve": 1e-6},
    },
}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nalgebra_kats.json")
with open(out, "w") as f:
    json.dump(kats, f, indent=1)
print("wrote", out)
