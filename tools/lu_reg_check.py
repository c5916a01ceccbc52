"""LU against the oracle (pivots bit-exact, packed factors) over shapes that exercise every GETF2 leaf
configuration (single CTA, multi-CTA, ragged widths, tall, wide, ties, zero columns, NaN), then panel
and full-matrix timings.  Usage: python tools/lu_reg_check.py [check|time|both]"""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import nalgebra_b200 as nab
from nalgebra_b200 import _capi
import oracle as O
eps = np.finfo(np.float64).eps
mode = sys.argv[1] if len(sys.argv) > 1 else "both"

def check(A, tag):
    m, n = A.shape
    lu = nab.LU.new(A)
    lur, swr = O.lu(A)
    same = np.array_equal(lu.p().ipiv, swr)
    a, b = lu.lu_internal(), lur
    eq = np.array_equal(a, b, equal_nan=True)
    err = np.nanmax(np.abs(a - b)) if a.size else 0.0
    print(f"lu {tag} {m}x{n}: pivots_equal={same} bit_equal={eq} max|lu-ref|={err:.2e}", flush=True)
    assert same, tag
    assert err < 1e-9 or np.isnan(err), (tag, err)
    return eq

if mode in ("check", "both"):
    shapes = [(1, 1), (2, 2), (3, 5), (5, 3), (17, 17), (40, 40), (64, 64), (65, 65), (100, 37), (37, 100), (128, 128), (129, 129),
              (130, 257), (257, 130), (300, 300), (513, 513), (640, 640), (1000, 1000), (1537, 1537), (2500, 48), (5000, 64),
              (9000, 33), (20000, 17), (3000, 3000)]
    alleq = True
    for (m, n) in shapes:
        alleq &= check(O.uniform(m, n, 6) - 0.3, "uniform")
    print("all bit-equal to the oracle:", alleq)
    # ties: entries from a tiny set of values, so every column has many equal |x| -> lowest index must win
    rng = np.random.default_rng(3)
    for (m, n) in [(50, 50), (700, 64), (1200, 200)]:
        check(rng.integers(-2, 3, size=(m, n)).astype(np.float64), "ties")
    # zero columns / zero pivots
    A = O.uniform(600, 80, 6) - 0.3; A[:, 3] = 0.0; A[:, 40] = 0.0; A[:, 79] = 0.0
    check(A, "zero-cols")
    A = np.zeros((300, 70)); check(A, "all-zero")
    # NaN: wins only at index 0 of the searched range
    A = O.uniform(400, 40, 6) - 0.3; A[5, 5] = np.nan; A[300, 20] = np.nan
    check(A, "nan")
    print("check ok", flush=True)

if mode in ("time", "both"):
    import torch
    L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    def run(M, N, reps=4):
        A0 = torch.empty(M * N, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
        _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), M, N, M, 6, s))
        swaps = (C.c_size_t * (2 * min(M, N)))(); ns = C.c_size_t(0)
        best = 1e9
        for it in range(reps):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            l0 = L.na_kernel_launches()
            e0.record(); _capi.check(L.na_lu_f64_dev(M, N, A.data_ptr(), M, swaps, C.addressof(ns), s)); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1)); nl = L.na_kernel_launches() - l0
        fl = (M * N * N - N ** 3 / 3.0) if M >= N else 0
        print(f"lu {M:6d} x {N:5d}: {best*1e3:9.1f} us  launches {nl}  ({best*1e3/min(M,N):6.2f} us/column, {fl/best/1e9:7.2f} TFLOP/s = {fl/best/1e9/37.18*100:5.1f} %)", flush=True)
    for (M, N) in [(256, 64), (512, 64), (2048, 64), (8192, 64), (16384, 64), (16384, 32), (16384, 16), (16384, 512), (8192, 512), (2048, 512),
                   (4096, 4096), (8192, 8192), (16384, 16384)]:
        run(M, N)
