"""Per-phase cycle counters of the register-resident GETF2 leaf (debug build:
NAB_EXTRA_NVCC_FLAGS="-DNAB_GETF2_PROF -DNAB_DEBUG_HOOKS" python -m nalgebra_b200.build, then
NAB_LIB=nalgebra_b200/libnalgebra_b200_dbg.so python tools/lu_reg_prof.py)."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
names = ["loop/other", "receive", "scale+col c+1", "candidate", "bulk", "publish_row"]
for (M, N) in [(256, 64), (2048, 64), (16384, 64), (16384, 32), (16384, 16), (60000, 32)]:
    A0 = torch.empty(M * N, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), M, N, M, 6, s))
    swaps = (C.c_size_t * (2 * N))(); ns = C.c_size_t(0)
    for it in range(2):
        A.copy_(A0); torch.cuda.synchronize()
        L.na_debug_getf2_reg_prof((C.c_longlong * 16)(), 1)
        _capi.check(L.na_lu_f64_dev(M, N, A.data_ptr(), M, swaps, C.addressof(ns), s)); torch.cuda.synchronize()
    prof = (C.c_longlong * 16)(); L.na_debug_getf2_reg_prof(prof, 0)
    tot = sum(prof[:6])
    print(f"{M} x {N}: total {tot} cycles = {tot/N:.0f} per column: " + ", ".join(f"{n} {prof[i]/N:.0f}" for i, n in enumerate(names)), flush=True)
