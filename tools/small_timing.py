"""Times small dependent launches back-to-back (real clocks, no profiler)."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
N = 16384
A = torch.empty(N * 2048, dtype=torch.float64, device=dev); B = torch.empty_like(A); Cd = torch.empty_like(A)
_capi.check(L.na_fill_uniform_dev(A.data_ptr(), N, 2048, N, 1, s)); _capi.check(L.na_fill_uniform_dev(B.data_ptr(), N, 2048, N, 2, s))
def timeit(name, fn, reps=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    import time; t0 = time.perf_counter()
    e0.record()
    for _ in range(reps): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name:55s} gpu {e0.elapsed_time(e1)/reps*1e3:8.1f} us/launch   host enqueue {(t1-t0)/reps*1e6:6.1f} us")
def gemm(m, k, n, beta=0.0, tb=False):
    rsb, csb = (N, 1) if tb else (1, N)
    return lambda: _capi.check(L.na_dgemm_dev(m, k, n, 1.0, A.data_ptr(), 1, N, B.data_ptr(), rsb, csb, beta, Cd.data_ptr(), 1, N, s))
for (m, k, n) in [(128,128,128),(16384,128,128),(16384,32,224)]:
    timeit(f"gemm m={m} k={k} n={n} beta=1", gemm(m, k, n, 1.0))
timeit("gemm 128x128x2048 B^T (row-major B)", gemm(128, 128, 2048, 0.0, True))
fail = C.c_size_t(0)
M = torch.eye(128, dtype=torch.float64, device=dev) * 4 + 0.01
def chol128(): _capi.check(L.na_cholesky_f64_dev(128, M.data_ptr(), 128, 0, 0.0, C.addressof(fail), s))
timeit("cholesky 128 (potf2_trtri + memset + sync)", chol128, 50)
M2 = torch.eye(64, dtype=torch.float64, device=dev) * 4 + 0.01
def chol64(): _capi.check(L.na_cholesky_f64_dev(64, M2.data_ptr(), 64, 0, 0.0, C.addressof(fail), s))
timeit("cholesky 64", chol64, 50)
M3 = torch.eye(8, dtype=torch.float64, device=dev) * 4 + 0.01
def chol8(): _capi.check(L.na_cholesky_f64_dev(8, M3.data_ptr(), 8, 0, 0.0, C.addressof(fail), s))
timeit("cholesky 8 (fixed overhead of the call)", chol8, 50)
T = torch.eye(1024, dtype=torch.float64, device=dev) * 4 + 0.01; X = torch.ones(1024 * 64, dtype=torch.float64, device=dev)
def trsm1024(): _capi.check(L.na_tri_solve_f64_dev(1, 0, 1, 1024, T.data_ptr(), 1024, X.data_ptr(), 1024, 64, s))
timeit("tri_solve unit lower 1024 x 64rhs (trtri 8 blocks + 15 gemm + 8 copy)", trsm1024, 20)
def trsm128(): _capi.check(L.na_tri_solve_f64_dev(1, 0, 1, 128, T.data_ptr(), 1024, X.data_ptr(), 1024, 64, s))
timeit("tri_solve unit lower 128 x 64rhs (trtri + gemm + copy)", trsm128, 50)
