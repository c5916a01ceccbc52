"""nalgebra_b200 -- B200-native (sm_100a) back end for nalgebra's DMatrix GEMM / Cholesky / LU / QR.

The product is ``libnalgebra_b200.so`` (C ABI: ``include/nalgebra_b200.h``); this package is the
host-side mirror of nalgebra's interface over that ABI (see ``linalg.py``).  No CPU fallback.
"""
from . import _capi  # noqa: F401
from . import wire  # noqa: F401
from .linalg import (  # noqa: F401
    bidiagonalize, cholesky, col_piv_qr, determinant, full_piv_lu, hessenberg, lu, qr, symmetric_tridiagonalize, try_inverse,
    LU, QR, Bidiagonal, Cholesky, ColPivQR, FullPivLU, Hessenberg, PermutationSequence, SymmetricTridiagonal, ad_mul, ad_mul_to, gemm, gemm_ad, gemm_f32, gemm_tr, gemv, gemv_ad, gemv_tr,
    kernel_launches, mul, mul_to, solve_lower_triangular, solve_lower_triangular_with_diag, solve_upper_triangular,
    syrk_lower, tr_mul, tr_mul_to, tr_solve_lower_triangular, tr_solve_upper_triangular, try_invert_to,
)
