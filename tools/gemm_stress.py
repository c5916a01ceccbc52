import sys
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
bad = 0
for (m, k, n, beta) in [(16384, 512, 15360, 1.0), (16384, 512, 15360, 0.0), (16384, 2048, 8192, 0.0)]:
    A = torch.rand(k, m, dtype=torch.float64, device=dev); B = torch.rand(n, k, dtype=torch.float64, device=dev); C0 = torch.rand(n, m, dtype=torch.float64, device=dev)
    ref = (beta * C0.t() - A.t() @ B.t())
    nb = 0; worst = 0.0
    for rep in range(reps):
        C = C0.clone()
        _capi.check(L.na_dgemm_dev(m, k, n, -1.0, A.data_ptr(), 1, m, B.data_ptr(), 1, k, beta, C.data_ptr(), 1, m, s))
        d = (C.t() - ref).abs()
        err = d.max().item()
        if err > 1e-9:
            nb += 1; worst = max(worst, err)
            idx = (d > 1e-9).nonzero()
            print("  rep", rep, "bad elements", idx.shape[0], "first", idx[0].tolist(), "last", idx[-1].tolist(), "err", err)
    print(f"m={m} k={k} n={n} beta={beta}: {nb}/{reps} bad runs, worst {worst}")
