"""Timeline of one host-pointer na_dgemm at 16384^3 (NAB_GEMM_TRACE=1) plus its wall time."""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); _capi.check(L.na_init(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
a = torch.rand(n * n, dtype=torch.float64).pin_memory(); b = torch.rand(n * n, dtype=torch.float64).pin_memory(); c = torch.empty(n * n, dtype=torch.float64).pin_memory()
for it in range(3):
    t0 = time.perf_counter()
    _capi.check(L.na_dgemm(n, n, n, 1.0, a.data_ptr(), 1, n, b.data_ptr(), 1, n, 0.0, c.data_ptr(), 1, n))
    print(f"== call {it}: wall {1e3*(time.perf_counter()-t0):.1f} ms", file=sys.stderr, flush=True)
