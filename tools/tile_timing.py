"""Per-tile cost of the DGEMM kernel at exact wave multiples (M = 148*128): mainloop vs epilogue."""
import sys
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
M = 148 * 128
A = torch.empty(M * 2048, dtype=torch.float64, device=dev); B = torch.empty(2048 * 2048, dtype=torch.float64, device=dev); Cd = torch.empty(M * 2048, dtype=torch.float64, device=dev)
_capi.check(L.na_fill_uniform_dev(A.data_ptr(), M, 2048, M, 1, s)); _capi.check(L.na_fill_uniform_dev(B.data_ptr(), 2048, 2048, 2048, 2, s))
def t(k, n, beta, reps=20):
    f = lambda: _capi.check(L.na_dgemm_dev(M, k, n, 1.0, A.data_ptr(), 1, M, B.data_ptr(), 1, 2048, beta, Cd.data_ptr(), 1, M, s))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    waves = n // 128
    ideal = 128 * 128 * k / 64 / 1965.0   # us per tile at the DMMA rate
    print(f"K={k:5d} N={n:5d} beta={beta}: {us:8.1f} us  = {us/waves:7.1f} us/tile-wave  (ideal mainloop {ideal:6.1f} us/tile, overhead {us/waves-ideal:6.1f} us, eff {ideal*waves/us*100:5.1f}%)")
for k in (128, 256, 512, 1024, 2048):
    for beta in (0.0, 1.0):
        t(k, 1024, beta)
