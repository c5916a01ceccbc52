import sys
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
for (m, k, n) in [(4096, 512, 4096), (16384, 512, 512), (8192, 128, 384), (3000, 512, 1000), (16384, 512, 15360)]:
    A = torch.rand(k, m, dtype=torch.float64, device=dev); B = torch.rand(n, k, dtype=torch.float64, device=dev); C0 = torch.rand(n, m, dtype=torch.float64, device=dev)
    # column-major views: A_cm (m x k) = A.t(), B_cm (k x n) = B.t(), C_cm (m x n) = C.t()
    for rep in range(3):
        C = C0.clone()
        _capi.check(L.na_dgemm_dev(m, k, n, -1.0, A.data_ptr(), 1, m, B.data_ptr(), 1, k, 1.0, C.data_ptr(), 1, m, s))
        ref = C0.t() - A.t() @ B.t()
        err = (C.t() - ref).abs().max().item()
        print(m, k, n, "beta=1 err", err)
    if m == n:
        C = C0.clone()
        _capi.check(L.na_dgemm_lower_dev(m, k, n, -1.0, A.data_ptr(), 1, m, A.data_ptr(), m, 1, 1.0, C.data_ptr(), m, s))
        ref = C0.t() - A.t() @ A
        err = (torch.tril(C.t() - ref)).abs().max().item(); up = (torch.triu(C.t() - C0.t(), 1)).abs().max().item()
        print(m, k, n, "lower-only err", err, "upper touched", up)
