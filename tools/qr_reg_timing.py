"""Register-resident GEQR2 leaf: per-phase cycle counters of CTA 1 (debug build:
NAB_EXTRA_NVCC_FLAGS="-DNAB_GEQR2_PROF -DNAB_DEBUG_HOOKS" python -m nalgebra_b200.build; NAB_LIB=nalgebra_b200/libnalgebra_b200_dbg.so)."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
for (M, N) in [(65536, 32), (16384, 32)]:
    A0 = torch.empty(M * N, dtype=torch.float64, device=dev); A = torch.empty_like(A0); d = torch.empty(N, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), M, N, M, 8, s))
    reps = 5; best = 1e9
    L.na_debug_geqr2r_prof((C.c_longlong * 16)(), 1)
    for it in range(reps):
        A.copy_(A0); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _capi.check(L.na_qr_f64_dev(M, N, A.data_ptr(), M, d.data_ptr(), s)); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    prof = (C.c_longlong * 16)(); L.na_debug_geqr2r_prof(prof, 1)
    cols = reps * N
    print(f"qr {M} x {N}: {best*1e3:8.1f} us ({best*1e3/N:6.2f} us/column)")
    print("   cycles/column [0 loop top, 1 receive, 2 reflector scalars + barriers, 3 pass arithmetic, 4 reduce-scatter, 5 barrier, 6 publish, 7 end barrier]:",
          [int(prof[i] / cols) for i in range(8)])
