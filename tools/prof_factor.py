"""One run of a factorization at N (device API) for ncu launch lists / captures.
Usage: python tools/prof_factor.py N chol|lu|hess|symtri|bidiag|fplu|cpqr    or    python tools/prof_factor.py M qr N"""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
N = int(sys.argv[1]); which = sys.argv[2]
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
A0 = torch.empty(N * N, dtype=torch.float64, device=dev)
if which == "chol":
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 5, s))
    M = A0.view(N, N); M.copy_((M + M.t()) * 0.5); M.diagonal().add_(float(N))
    fail = C.c_size_t(0)
    torch.cuda.synchronize()
    import time; t0 = time.perf_counter()
    print(L.na_cholesky_f64_dev(N, A0.data_ptr(), N, 0, 0.0, C.addressof(fail), s), "host wall ms", (time.perf_counter() - t0) * 1e3)
elif which == "lu":
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 6, s))
    swaps = (C.c_size_t * (2 * N))(); ns = C.c_size_t(0)
    torch.cuda.synchronize()
    print(L.na_lu_f64_dev(N, N, A0.data_ptr(), N, swaps, C.addressof(ns), s))
elif which == "qr":
    m, n = N, int(sys.argv[3])
    A0 = torch.empty(m * n, dtype=torch.float64, device=dev); d = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), m, n, m, 8, s))
    torch.cuda.synchronize()
    print(L.na_qr_f64_dev(m, n, A0.data_ptr(), m, d.data_ptr(), s))
elif which in ("hess", "symtri", "bidiag", "fplu", "cpqr"):
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 6, s))
    d = torch.empty(N, dtype=torch.float64, device=dev); e = torch.empty(N, dtype=torch.float64, device=dev)
    ps = (C.c_size_t * (2 * N))(); qs = (C.c_size_t * (2 * N))(); a_ = C.c_size_t(0); b_ = C.c_size_t(0)
    torch.cuda.synchronize()
    if which == "hess": print(L.na_hessenberg_f64_dev(N, A0.data_ptr(), N, d.data_ptr(), s))
    elif which == "symtri": print(L.na_symmetric_tridiagonal_f64_dev(N, A0.data_ptr(), N, d.data_ptr(), s))
    elif which == "bidiag": print(L.na_bidiagonal_f64_dev(N, N, A0.data_ptr(), N, d.data_ptr(), e.data_ptr(), s))
    elif which == "fplu": print(L.na_full_piv_lu_f64_dev(N, N, A0.data_ptr(), N, ps, C.addressof(a_), qs, C.addressof(b_), s))
    else: print(L.na_col_piv_qr_f64_dev(N, N, A0.data_ptr(), N, d.data_ptr(), ps, C.addressof(a_), s))
torch.cuda.synchronize()
