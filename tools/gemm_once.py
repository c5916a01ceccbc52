"""One device-resident 16384^3 f64 GEMM after a warm-up (for `ncu --set full -k regex:dgemm_tma -c 1`). Usage: python tools/gemm_once.py [N]"""
import sys
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A = torch.empty(N * N, dtype=torch.float64, device=dev); B = torch.empty_like(A); C = torch.empty_like(A)
_capi.check(L.na_fill_uniform_dev(A.data_ptr(), N, N, N, 1, s)); _capi.check(L.na_fill_uniform_dev(B.data_ptr(), N, N, N, 2, s))
for _ in range(2):
    _capi.check(L.na_dgemm_dev(N, N, N, 1.0, A.data_ptr(), 1, N, B.data_ptr(), 1, N, 0.0, C.data_ptr(), 1, N, s))
torch.cuda.synchronize()
