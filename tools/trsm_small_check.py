"""Isolated check of trsm_unit_lower_small_kernel against numpy."""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from nalgebra_b200 import _capi
L = _capi.lib(); _capi.check(L.na_init(0))
f = L.na_debug_trsm_unit_lower_small
f.restype = C.c_int; f.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
s = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(1)
for n1 in (64, 128, 100, 37, 65, 127):
    for nrhs in (1, 3, 64, 65, 200):
        ldl = n1 + 6; ldb = n1 + 10
        Lm = np.asfortranarray(np.tril(rng.random((ldl, n1)) - 0.5, -1)); Lfull = Lm.copy(); Lfull[:n1][np.triu_indices(n1)] = 7.7   # garbage on/above diag
        B = np.asfortranarray(rng.random((ldb, nrhs)))
        ref = np.linalg.solve(np.tril(Lm[:n1], -1) + np.eye(n1), B[:n1])
        dl = torch.from_numpy(Lfull.T.copy()).cuda(); db = torch.from_numpy(B.T.copy()).cuda()     # (cols, ld) row-major == column-major buffer
        _capi.check(f(n1, dl.data_ptr(), ldl, db.data_ptr(), ldb, nrhs, s)); torch.cuda.synchronize()
        got = db.cpu().numpy().T
        err = np.abs(got[:n1] - ref).max(); pad = np.abs(got[n1:] - B[n1:]).max()
        bad = np.argwhere(np.abs(got[:n1] - ref) > 1e-9)
        print(f"n1={n1:4d} nrhs={nrhs:4d}: err {err:.2e} pad-touched {pad:.1e}", ("first bad (row, col) %s, bad rows %s" % (bad[0].tolist(), sorted(set(bad[:, 0].tolist()))[:12])) if len(bad) else "")
