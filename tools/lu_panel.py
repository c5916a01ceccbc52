"""LU of tall panels (the lu_rec path of one outer block) and of the full matrix: timing + launch counts.
Usage: python tools/lu_panel.py [once M N]   ('once' = single run, for ncu launch lists)"""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
def run(M, N, reps=5):
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
    print(f"lu {M:6d} x {N:5d}: {best*1e3:9.1f} us  launches {nl}  ({best*1e3/min(M,N):6.2f} us/column, {fl/best/1e9:7.2f} TFLOP/s)", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "once":
    run(int(sys.argv[2]), int(sys.argv[3]), reps=1)
else:
    for (M, N) in [(16384, 64), (16384, 512), (8192, 512), (2048, 512), (4096, 4096), (8192, 8192), (16384, 16384)]:
        run(M, N)
