"""Per-phase time of the Hessenberg / SymmetricTridiagonal kernel (debug build: NAB_EXTRA_NVCC_FLAGS="-DNAB_TS_PROF
-DNAB_DEBUG_HOOKS" python -m nalgebra_b200.build; NAB_LIB=nalgebra_b200/libnalgebra_b200_dbg.so python tools/twosided_prof.py [n ...])"""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
NAMES = ["tiling", "pass1", "B1 wait", "pass2/col", "axis", "B2 wait"]
for n in [int(x) for x in sys.argv[1:]] or [1024, 4096]:
    A0 = torch.empty(n * n, dtype=torch.float64, device=dev); A = torch.empty_like(A0); d = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), n, n, n, 6, s))
    out = (C.c_ulonglong * 16)()
    for name, fn in [("hessenberg", lambda: L.na_hessenberg_f64_dev(n, A.data_ptr(), n, d.data_ptr(), s)),
                     ("symmetric_tridiagonal", lambda: L.na_symmetric_tridiagonal_f64_dev(n, A.data_ptr(), n, d.data_ptr(), s))]:
        for rep in range(2):
            A.copy_(A0); torch.cuda.synchronize(); L.na_debug_ts_prof(out)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); _capi.check(fn()); e1.record(); torch.cuda.synchronize()
            L.na_debug_ts_prof(out)
        print(f"{name} n={n}: {e0.elapsed_time(e1):.2f} ms; us/step by phase:")
        for who, off in (("task CTA 0", 0), ("leader", 8)):
            print("   ", who, "  ".join(f"{NAMES[i]} {out[off + i] / 1e3 / n:.2f}" for i in range(6)), flush=True)
