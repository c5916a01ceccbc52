"""Device-resident timing of Cholesky / LU (N x N) and QR (M x N): python tools/factor_timing.py chol,lu,qr [N]"""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
which = sys.argv[1] if len(sys.argv) > 1 else "chol,lu"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
PEAK = 37.18e12
def timed(f, setup, reps=4):
    best = 1e9
    for _ in range(reps):
        setup(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
if "chol" in which:
    A0 = torch.empty(N * N, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
    _capi.check(L.na_fill_spd_block_dev(A0.data_ptr(), N, N, N, 5, 0, 0, N, s))
    fail = C.c_size_t(0)
    ms = timed(lambda: _capi.check(L.na_cholesky_f64_dev(N, A.data_ptr(), N, 0, 0.0, C.addressof(fail), s)), lambda: A.copy_(A0))
    fl = N ** 3 / 3.0
    print(f"cholesky N={N}: {ms:8.2f} ms  {fl/ms/1e9:7.2f} TFLOP/s  {100*fl/ms/1e-3/PEAK:5.1f}% of peak", flush=True)
    del A0, A
if "lu" in which:
    A0 = torch.empty(N * N, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 6, s))
    swaps = (C.c_size_t * (2 * N))(); ns = C.c_size_t(0)
    ms = timed(lambda: _capi.check(L.na_lu_f64_dev(N, N, A.data_ptr(), N, swaps, C.addressof(ns), s)), lambda: A.copy_(A0))
    fl = 2 * N ** 3 / 3.0
    print(f"lu       N={N}: {ms:8.2f} ms  {fl/ms/1e9:7.2f} TFLOP/s  {100*fl/ms/1e-3/PEAK:5.1f}% of peak  nswaps {ns.value}", flush=True)
    del A0, A
if "qr" in which:
    m, n = 65536, 4096
    A0 = torch.empty(m * n, dtype=torch.float64, device=dev); A = torch.empty_like(A0); d = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), m, n, m, 8, s))
    ms = timed(lambda: _capi.check(L.na_qr_f64_dev(m, n, A.data_ptr(), m, d.data_ptr(), s)), lambda: A.copy_(A0))
    fl = 2.0 * m * n * n - 2.0 * n ** 3 / 3.0
    print(f"qr {m}x{n}: {ms:8.2f} ms  {fl/ms/1e9:7.2f} TFLOP/s  {100*fl/ms/1e-3/PEAK:5.1f}% of peak", flush=True)
