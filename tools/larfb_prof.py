"""Per-phase cycle counters of the fused in-panel block reflector (CTA 1), debug build:
NAB_EXTRA_NVCC_FLAGS="-DNAB_LARFB_PROF -DNAB_DEBUG_HOOKS" python -m nalgebra_b200.build; NAB_LIB=nalgebra_b200/libnalgebra_b200_dbg.so python tools/larfb_prof.py"""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
m = 65536
for n in [64, 128, 256]:
    A0 = torch.empty(m * n, dtype=torch.float64, device=dev); A = torch.empty_like(A0); d = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), m, n, m, 8, s))
    for it in range(3):
        A.copy_(A0); torch.cuda.synchronize()
        L.na_debug_larfb_prof((C.c_longlong * 16)(), 1)
        _capi.check(L.na_qr_f64_dev(m, n, A.data_ptr(), m, d.data_ptr(), s)); torch.cuda.synchronize()
    prof = (C.c_longlong * 16)(); L.na_debug_larfb_prof(prof, 1)
    leaves = n // 32 - 1
    print(f"qr {m} x {n}: {leaves} fused leaves; cycles per leaf [phase 1, hand-off 1 + reduce, hand-off 2 + totals, S + solve, phase 4]:",
          [int(prof[i] / max(1, leaves)) for i in range(5)])
