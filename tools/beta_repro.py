import sys
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
n = 16384
A = torch.empty(n * n, dtype=torch.float64, device=dev); B = torch.empty_like(A); Cd = torch.empty_like(A)
_capi.check(L.na_fill_uniform_dev(A.data_ptr(), n, n, n, 1, s)); _capi.check(L.na_fill_uniform_dev(B.data_ptr(), n, n, n, 2, s))
_capi.check(L.na_dgemm_dev(n, n, n, 1.0, A.data_ptr(), 1, n, B.data_ptr(), 1, n, 0.0, Cd.data_ptr(), 1, n, s))
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    C2 = Cd.clone()
    _capi.check(L.na_dgemm_dev(n, n, n, 0.5, A.data_ptr(), 1, n, B.data_ptr(), 1, n, 0.5, C2.data_ptr(), 1, n, s))
    d = (C2 - Cd).abs().view(n, n)          # d[col, row] (column-major buffer)
    bad = (d > 1e-6).nonzero()
    if bad.shape[0] == 0:
        print("rep", rep, "ok"); continue
    cols, rows = bad[:, 0], bad[:, 1]
    print("rep", rep, "bad", bad.shape[0], "rows", rows.min().item(), rows.max().item(), "cols", cols.min().item(), cols.max().item(),
          "tiles(m,n):", sorted(set(zip((rows // 128).tolist(), (cols // 128).tolist())))[:6], "max", d.max().item())
    r0 = (rows.min().item() // 128) * 128; c0 = (cols.min().item() // 128) * 128
    sub = (d[c0:c0 + 128, r0:r0 + 128] > 1e-6)
    print("   within first bad tile: bad rows(local)", sorted(set(sub.nonzero()[:, 1].tolist()))[:20], "bad cols(local)", sorted(set(sub.nonzero()[:, 0].tolist()))[:20])
    vals = (C2.view(n, n)[c0:c0 + 128, r0:r0 + 128] - Cd.view(n, n)[c0:c0 + 128, r0:r0 + 128])[sub]
    print("   sample diffs", vals[:6].tolist(), " Cd there", Cd.view(n, n)[c0:c0 + 128, r0:r0 + 128][sub][:3].tolist())
