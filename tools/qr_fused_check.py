"""Fused in-panel block reflector (panel_qr_fused.cu, leaves with >= 2048 rows) against the oracle on shapes that
exercise ragged leaves (n % 32 != 0, n % 8 != 0), then timing at 65536 x 4096 with the fused path on and off."""
import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import nalgebra_b200 as nab
import oracle as O
eps = np.finfo(np.float64).eps
for (m, n) in [(2100, 70), (2500, 300), (3000, 333), (4100, 257), (8300, 1100)]:
    A = O.uniform(m, n, 8) - 0.4
    if n == 333: A[:, 40] = 0.0                       # a zero column inside a fused leaf (tau = 0)
    qr = nab.QR.new(A)
    qr_ref, diag_ref = O.qr(A)
    e1 = np.abs(qr.qr_internal() - qr_ref).max(); e2 = np.abs(qr.diag_internal() - diag_ref).max()
    print(f"qr {m}x{n}: |qr-ref|={e1:.2e} |diag-ref|={e2:.2e}", flush=True)
    assert e1 < 1e-9 and e2 < 1e-9
print("qr fused ok")
