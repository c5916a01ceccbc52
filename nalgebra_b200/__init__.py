"""nalgebra_b200 -- B200-native (sm_100a) back end for nalgebra's DMatrix GEMM / Cholesky / LU / QR.

The product is ``libnalgebra_b200.so`` (C ABI: ``include/nalgebra_b200.h``); this package is the
host-side mirror of nalgebra's interface over that ABI (see ``linalg.py``).  No CPU fallback.
"""
from . import _capi  # noqa: F401
from .linalg import (  # noqa: F401
    LU, QR, Cholesky, PermutationSequence, gemm, gemm_f32, gemm_tr, kernel_launches, mul, mul_to,
    solve_lower_triangular, solve_lower_triangular_with_diag, solve_upper_triangular, tr_mul,
    tr_solve_lower_triangular, tr_solve_upper_triangular,
)
