"""Device-resident timing of Hessenberg / SymmetricTridiagonal / Bidiagonal (n x n): python tools/twosided_timing.py [n ...]
Algorithmic traffic per step: one read + one read-modify-write of the trailing block --
Hessenberg 24 n (n - k) bytes, SymmetricTridiagonal 12 (n - k)^2, Bidiagonal 32 (n - k)^2."""
import sys
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
for n in [int(x) for x in sys.argv[1:]] or [1024, 2048, 4096]:
    A0 = torch.empty(n * n, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
    d = torch.empty(n, dtype=torch.float64, device=dev); e = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), n, n, n, 6, s))
    for name, fn, nbytes in [("hessenberg", lambda: L.na_hessenberg_f64_dev(n, A.data_ptr(), n, d.data_ptr(), s), 12.0 * n ** 3),
                             ("symmetric_tridiagonal", lambda: L.na_symmetric_tridiagonal_f64_dev(n, A.data_ptr(), n, d.data_ptr(), s), 4.0 * n ** 3),
                             ("bidiagonal", lambda: L.na_bidiagonal_f64_dev(n, n, A.data_ptr(), n, d.data_ptr(), e.data_ptr(), s), 32.0 * n ** 3 / 3)]:
        best = 1e9
        for _ in range(2):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); _capi.check(fn()); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name} n={n}: {best:8.2f} ms  {nbytes / best / 1e6:7.1f} GB/s algorithmic  ({best * 1e3 / n:.1f} us/step)", flush=True)
