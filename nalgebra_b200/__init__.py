"""nalgebra_b200 -- B200-native (sm_100a) back end for nalgebra's DMatrix GEMM / Cholesky / LU / QR."""
from . import _capi  # noqa: F401
