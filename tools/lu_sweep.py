"""LU N x N timing for the current environment knobs (NAB_LU_*). Usage: python tools/lu_sweep.py [N] [reps]"""
import sys, os, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
A0 = torch.empty(N * N, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
_capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 6, s))
swaps = (C.c_size_t * (2 * N))(); ns = C.c_size_t(0)
best = 1e9
for it in range(reps + 1):
    A.copy_(A0); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); _capi.check(L.na_lu_f64_dev(N, N, A.data_ptr(), N, swaps, C.addressof(ns), s)); e1.record(); torch.cuda.synchronize()
    if it: best = min(best, e0.elapsed_time(e1))
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("NAB_"))
print(f"lu {N}: {best:8.2f} ms  {2*N**3/3/best/1e9:6.2f} TFLOP/s = {2*N**3/3/best/1e9/37.18*100:5.1f} %   [{knobs}]", flush=True)
