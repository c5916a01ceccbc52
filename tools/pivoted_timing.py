"""Device-resident timing of FullPivLU / ColPivQR (n x n): python tools/pivoted_timing.py [n ...]"""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
for n in [int(x) for x in sys.argv[1:]] or [1024, 2048, 4096]:
    A0 = torch.empty(n * n, dtype=torch.float64, device=dev); A = torch.empty_like(A0); d = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), n, n, n, 6, s))
    ps = np.zeros(2 * n, dtype=np.uint64); qs = np.zeros(2 * n, dtype=np.uint64); a_, b_ = C.c_size_t(0), C.c_size_t(0)
    for name, fn in [("full_piv_lu", lambda: L.na_full_piv_lu_f64_dev(n, n, A.data_ptr(), n, ps.ctypes.data, C.addressof(a_), qs.ctypes.data, C.addressof(b_), s)),
                     ("col_piv_qr", lambda: L.na_col_piv_qr_f64_dev(n, n, A.data_ptr(), n, d.data_ptr(), ps.ctypes.data, C.addressof(a_), s))]:
        best = 1e9
        for _ in range(2):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); _capi.check(fn()); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        # one read + one write of the trailing matrix per step: 16 * n^3 / 3 bytes
        print(f"{name} n={n}: {best:8.2f} ms  {16 * n ** 3 / 3 / best / 1e6:7.1f} GB/s of trailing-matrix traffic  ({best * 1e3 / n:.1f} us/step)", flush=True)
