import sys, ctypes as C, time
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
def run(M, N, reps=10):
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
    print(f"lu {M:6d} x {N:4d}: {best*1e3:9.1f} us  launches {nl}  ({best*1e3/min(M,N):6.2f} us/column)")
    prof = (C.c_longlong * 16)(); L.na_debug_getf2_prof(prof, 1)
    cols = reps * min(M, N)
    print("     cycles/column [top, 1 poll headers, 1 reduce, 1 rows, 2 swap+scale, 3 col c+1 + publish, 4 bulk update]:", [int(prof[i] / cols) for i in range(7)])
L.na_debug_getf2_prof((C.c_longlong * 16)(), 1)
for (M, N) in [(128, 128), (1024, 128), (16384, 128), (16384, 32)]:
    run(M, N)
