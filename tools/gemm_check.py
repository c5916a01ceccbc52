"""Quick GPU check of na_dgemm (host + device API): correctness on assorted shapes/strides, then
timing at a large size.  Usage: python tools/gemm_check.py [N]"""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
from nalgebra_b200 import _capi
import oracle as O
L = _capi.lib()
print(L.na_version().decode())
_capi.check(L.na_init(0))

def st(a):
    return a.strides[0] // 8, (a.strides[1] // 8 if a.ndim > 1 else 0)

def gemm_host(alpha, a, b, beta, c):
    m, k = a.shape; n = b.shape[1]
    _capi.check(L.na_dgemm(m, k, n, alpha, a.ctypes.data, *st(a), b.ctypes.data, *st(b), beta, c.ctypes.data, *st(c)))

rng = np.random.default_rng(0)
worst = 0
for (m, k, n) in [(1,1,1),(2,3,4),(7,5,3),(8,8,8),(16,16,16),(17,33,9),(64,64,64),(127,129,130),(128,16,128),(200,1000,50),(513,257,255),(1024,1024,1024)]:
    for ta in (0, 1):
        for tb in (0, 1):
            for tc in (0, 1):
                for (alpha, beta) in [(1.0, 0.0), (1.5, 0.5)]:
                    a = O.uniform(m, k, 1) - 0.5; b = O.uniform(k, n, 2) - 0.5; c0 = O.uniform(m, n, 3)
                    A = np.ascontiguousarray(a) if ta else np.asfortranarray(a)
                    B = np.ascontiguousarray(b) if tb else np.asfortranarray(b)
                    Cm = c0.copy(order='C') if tc else c0.copy(order='F')
                    if beta == 0.0: Cm[:] = np.nan
                    gemm_host(alpha, A, B, beta, Cm)
                    ref = alpha * (a @ b) + (beta * c0 if beta else 0)
                    err = np.abs(Cm - ref).max()
                    tol = 4 * k * 2.2e-16 * np.abs(a).max() * np.abs(b).max() * k + 1e-300
                    worst = max(worst, err / tol)
                    if not err <= tol:
                        print("FAIL", m, k, n, ta, tb, tc, alpha, beta, err, tol); sys.exit(1)
print("host-API correctness ok; worst err/tol", worst)
# strided views (both strides != 1) and odd leading dims
big = O.uniform(300, 300, 5)
a = big[3:150:2, 1:200:3]; b = big[5:5+a.shape[1], 7:90]; c = np.zeros((a.shape[0], b.shape[1]))[:, :]
cv = np.zeros((2*a.shape[0], 3*b.shape[1]))[::2, ::3]
gemm_host(1.0, a, b, 0.0, cv)
print("strided view err", np.abs(cv - a @ b).max())
# k == 0
c = np.ones((13, 14), order="F"); gemm_host(1.0, np.zeros((13, 0), order="F"), np.zeros((0, 14), order="F"), 0.5, c); assert (c == 0.5).all()
c = np.full((13, 14), np.nan, order="F"); gemm_host(1.0, np.zeros((13, 0), order="F"), np.zeros((0, 14), order="F"), 0.0, c); assert (c == 0).all()
print("k==0 ok")

# ---- timing, device API ----
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
import torch
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
A = torch.empty((N, N), dtype=torch.float64, device=dev); B = torch.empty_like(A); Cd = torch.empty_like(A)
_capi.check(L.na_fill_uniform_dev(A.data_ptr(), N, N, N, 1, s)); _capi.check(L.na_fill_uniform_dev(B.data_ptr(), N, N, N, 2, s))
def run(ta=False, tb=False, beta=0.0):
    # torch tensors are row-major: treat the buffer as column-major N x N (ld = N)
    rsa, csa = (N, 1) if ta else (1, N); rsb, csb = (N, 1) if tb else (1, N)
    _capi.check(L.na_dgemm_dev(N, N, N, 1.0, A.data_ptr(), rsa, csa, B.data_ptr(), rsb, csb, beta, Cd.data_ptr(), 1, N, s))
for (ta, tb, beta) in [(False, False, 0.0), (False, True, 0.0), (True, False, 0.0), (True, True, 0.0), (False, False, 0.5)]:
    run(ta, tb, beta); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps): run(ta, tb, beta)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"N={N} ta={ta} tb={tb} beta={beta}: {ms:.2f} ms  {2*N**3/ms/1e9:.2f} TFLOP/s  ({2*N**3/ms/1e9/37.18*100:.1f}% of 37.18)")
# spot check the device result against torch matmul on a slice
run(False, False, 0.0); torch.cuda.synchronize()
Am = A.t()[:256, :]; Bm = B.t()[:, :256]   # column-major interpretation
ref = Am @ Bm
got = Cd.t()[:256, :256]
print("device spot-check max err", (ref - got).abs().max().item())
