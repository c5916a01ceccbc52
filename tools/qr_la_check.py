"""QR look-ahead path (m >= 8192, k >= 1024) against the oracle at a size the oracle finishes in seconds, and timing at 65536x4096."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import nalgebra_b200 as nab
import oracle as O
eps = np.finfo(np.float64).eps
m, n = 8300, 1100
A = O.uniform(m, n, 8) - 0.4
t0 = time.time(); qr = nab.QR.new(A); t1 = time.time()
qr_ref, diag_ref = O.qr(A); t2 = time.time()
print(f"gpu {t1-t0:.2f}s oracle {t2-t1:.2f}s")
e1 = np.abs(qr.qr_internal() - qr_ref).max(); e2 = np.abs(qr.diag_internal() - diag_ref).max()
q = qr.q(); r = qr.r()
res = np.linalg.norm(q @ r - A) / np.linalg.norm(A); orth = np.abs(q.T @ q - np.eye(n)).max()
print(f"qr {m}x{n} (look-ahead): |qr-ref|={e1:.2e} |diag-ref|={e2:.2e} resid={res:.2e} orth={orth:.2e}")
assert e1 < 1e-9 and e2 < 1e-9 and res <= 10 * m * eps and orth <= 10 * m * eps
print("qr look-ahead ok")
