"""Host-pointer (pinned) Cholesky / QR against the device-resident calls: same bits, wall time next to device time.
Usage: python tools/e2e_factor.py [chol,qr] [N]"""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
which = sys.argv[1] if len(sys.argv) > 1 else "chol,qr"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
if "chol" in which:
    A0 = torch.empty(N * N, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_spd_block_dev(A0.data_ptr(), N, N, N, 5, 0, 0, N, s))
    A = A0.clone(); fail = C.c_size_t(0)
    _capi.check(L.na_cholesky_f64_dev(N, A.data_ptr(), N, 0, 0.0, C.addressof(fail), s)); torch.cuda.synchronize()
    hA0 = A0.cpu(); hA = torch.empty(N * N, dtype=torch.float64).pin_memory()
    for it in range(3):
        hA.copy_(hA0)
        t0 = time.perf_counter(); st = L.na_cholesky_f64(N, hA.data_ptr(), N, 0, 0.0, C.addressof(fail)); dt = (time.perf_counter() - t0) * 1e3
        print(f"na_cholesky_f64 N={N}: status {st} {dt:.1f} ms", flush=True)
    G = A.cpu().view(N, N).t(); H = hA.view(N, N).t()
    lower_equal = bool(torch.equal(torch.tril(G), torch.tril(H)))
    upper_untouched = bool(torch.equal(torch.triu(H, 1), torch.triu(hA0.view(N, N).t(), 1)))
    print("  lower triangle == device-resident result:", lower_equal, " strict upper untouched on the host:", upper_untouched)
    assert lower_equal and upper_untouched
    # not positive definite: status and failing column as before
    hA.copy_(hA0); hA.view(N, N)[N // 2, N // 2] = -1.0
    st = L.na_cholesky_f64(N, hA.data_ptr(), N, 0, 0.0, C.addressof(fail)); print("  not-PD status", st, "fail column", fail.value)
    assert st == 1 and fail.value == N // 2
    del A0, A, hA, hA0
if "qr" in which:
    m, n = (65536, 4096) if N >= 16384 else (8300, 1100)
    A0 = torch.empty(m * n, dtype=torch.float64, device=dev); d = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), m, n, m, 8, s))
    A = A0.clone()
    _capi.check(L.na_qr_f64_dev(m, n, A.data_ptr(), m, d.data_ptr(), s)); torch.cuda.synchronize()
    hA0 = A0.cpu(); hA = torch.empty(m * n, dtype=torch.float64).pin_memory(); hd = torch.empty(n, dtype=torch.float64)
    for it in range(3):
        hA.copy_(hA0)
        t0 = time.perf_counter(); st = L.na_qr_f64(m, n, hA.data_ptr(), m, hd.data_ptr()); dt = (time.perf_counter() - t0) * 1e3
        print(f"na_qr_f64 {m}x{n}: status {st} {dt:.1f} ms", flush=True)
    eq = bool(torch.equal(A.cpu(), hA)) and bool(torch.equal(d.cpu(), hd))
    print("  storage and diag == device-resident result:", eq)
    assert eq
