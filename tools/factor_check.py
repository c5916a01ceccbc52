"""GPU check of Cholesky / LU / solves against the oracle. Usage: python tools/factor_check.py [bigN]"""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import nalgebra_b200 as nab
from nalgebra_b200 import _capi
import oracle as O
eps = np.finfo(np.float64).eps
which = sys.argv[2] if len(sys.argv) > 2 else "chol,lu,qr"

if "chol" in which and "cholbig" not in which:
    for n in [1, 2, 5, 17, 64, 128, 129, 200, 257, 640, 1000, 1500]:
        A = O.spd_wellcond(n, 5)
        A[0, n - 1] = np.nan if n > 1 else A[0, 0]        # strict upper is never read
        ch = nab.Cholesky.new(A)
        assert ch is not None, n
        Lref = np.tril(O.cholesky(A))
        L = ch.l()
        As = np.tril(A) + np.tril(A, -1).T
        res = np.linalg.norm(L @ L.T - As) / np.linalg.norm(As)
        assert res <= 10 * n * eps, (n, res)
        assert np.abs(L - Lref).max() <= 100 * n * eps * np.abs(Lref).max(), (n, np.abs(L - Lref).max())
        if n > 1: assert np.isnan(ch.chol[0, n - 1])      # upper untouched
        b = O.uniform(n, 3, 9)
        x = ch.solve(b); xr = O.cholesky_solve(O.cholesky(A), b)
        assert np.abs(x - xr).max() <= 1e-9 * max(1, np.abs(xr).max()), (n, np.abs(x - xr).max())
    # not PD
    A = O.spd_wellcond(300, 5); A[150, 150] = -1.0
    assert nab.Cholesky.new(A) is None and O.cholesky(A) is None
    m = np.array([[1.0, np.nan], [1.0, 1e-32]])
    assert nab.Cholesky.new(m) is None and nab.Cholesky.new_with_substitute(m, 1e-8) is not None
    print("cholesky ok")

if "lu" in which:
    for (m, n) in [(1, 1), (2, 2), (3, 5), (5, 3), (17, 17), (64, 64), (128, 128), (129, 129), (130, 257), (257, 130), (300, 300), (640, 640), (1000, 1000), (1537, 1537)]:
        A = O.uniform(m, n, 6) - 0.3
        lu = nab.LU.new(A)
        lur, swr = O.lu(A)
        same = np.array_equal(lu.p().ipiv, swr)
        err = np.abs(lu.lu_internal() - lur).max()
        l, u = lu.l(), lu.u()
        rec = l @ u; lu.p().inv_permute_rows(rec)
        res = np.linalg.norm(rec - A) / np.linalg.norm(A)
        print(f"lu {m}x{n}: pivots_equal={same} max|lu-lu_ref|={err:.2e} resid={res:.2e}")
        assert same and res <= 10 * max(m, n) * eps and err < 1e-9
    A = O.uniform(500, 500, 6); lu = nab.LU.new(A); b = O.uniform(500, 7, 7)
    x = lu.solve(b); lur, swr = O.lu(A); xr = O.lu_solve(lur, swr, b)
    print("lu solve err vs oracle", np.abs(x - xr).max(), "resid", np.abs(A @ x - b).max())
    assert np.abs(x - xr).max() < 1e-8
    # zero column & singular
    A = O.uniform(50, 50, 6); A[:, 3] = 0.0
    lu = nab.LU.new(A); lur, swr = O.lu(A)
    assert np.array_equal(lu.p().ipiv, swr) and np.abs(lu.lu_internal() - lur).max() < 1e-10
    assert lu.solve(np.ones((50, 1))) is None and not lu.is_invertible()
    m3 = np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]]); assert nab.LU.new(m3).determinant() == 4.0
    m3 = np.array([[0.0, -1, 2], [-1, 2, -1], [2, -1, 0]]); assert nab.LU.new(m3).determinant() == -4.0
    print("lu ok")

if "qr" in which:
    for (m, n) in [(1, 1), (3, 1), (1, 4), (7, 5), (5, 7), (33, 33), (40, 40), (64, 20), (100, 64), (257, 130), (130, 257), (300, 300), (700, 300), (1000, 513), (2000, 600)]:
        A = O.uniform(m, n, 8) - 0.5
        qr = nab.QR.new(A)
        qref, dref = O.qr(A)
        k = min(m, n)
        e1 = np.abs(qr.qr_internal() - qref).max(); e2 = np.abs(qr.diag_internal() - dref).max()
        q, r = qr.q(), qr.r()
        res = np.linalg.norm(q @ r - A) / np.linalg.norm(A); orth = np.linalg.norm(q.T @ q - np.eye(k))
        qo = O.qr_q(qref, dref)
        B = O.uniform(m, 3, 4); Bq = B.copy(order="F"); qr.q_tr_mul(Bq)
        e3 = np.abs(Bq - O.qr_q_tr_mul(qref, dref, B)).max()
        print(f"qr {m}x{n}: |qr-ref|={e1:.2e} |diag-ref|={e2:.2e} |q-qref|={np.abs(q-qo).max():.2e} resid={res:.2e} orth={orth:.2e} qtmul={e3:.2e}")
        assert e1 < 1e-10 and e2 < 1e-10 and res <= 10 * max(m, n) * eps and orth <= 10 * max(m, n) * eps and e3 < 1e-10
    A = O.uniform(200, 200, 8); qr = nab.QR.new(A); b = O.uniform(200, 5, 7); x = qr.solve(b)
    print("qr solve resid", np.abs(A @ x - b).max()); assert np.abs(A @ x - b).max() < 1e-9
    A = O.uniform(20, 12, 8) - 0.5; A[:, 4] = 0.0
    qr = nab.QR.new(A); qref, dref = O.qr(A)
    print("qr zero col", np.abs(qr.qr_internal() - qref).max(), np.abs(qr.diag_internal() - dref).max())
    assert np.abs(qr.qr_internal() - qref).max() < 1e-12
    print("qr ok")
    a32 = (O.uniform(300, 200, 1) - 0.5).astype(np.float32); b32 = (O.uniform(200, 150, 2) - 0.5).astype(np.float32)
    c32 = np.ones((300, 150), dtype=np.float32, order="F")
    nab.gemm_f32(1.5, a32, b32.T.copy().T, 0.5, c32)
    print("sgemm err", np.abs(c32 - (1.5 * a32.astype(np.float64) @ b32.astype(np.float64) + 0.5)).max())

N = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if N:
    import torch
    L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    if "chol" in which or "cholbig" in which:
        A0 = torch.empty(N * N, dtype=torch.float64, device=dev)
        _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 5, s))
        M = A0.view(N, N); M.copy_((M + M.t()) * 0.5); M.diagonal().add_(float(N))
        A = torch.empty_like(A0); fail = C.c_size_t(0)
        for it in range(3):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            l0 = L.na_kernel_launches()
            e0.record(); st = L.na_cholesky_f64_dev(N, A.data_ptr(), N, 0, 0.0, C.addressof(fail), s); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"cholesky N={N}: status {st} {ms:.2f} ms {N**3/3/ms/1e9:.2f} TFLOP/s ({N**3/3/ms/1e9/37.18*100:.1f}%) launches {L.na_kernel_launches()-l0}")
        Lm = torch.tril(A.view(N, N).t())            # column-major buffer viewed by torch as its transpose
        R = Lm @ Lm.t() - A0.view(N, N).t()
        print("  resid", (torch.linalg.norm(torch.tril(R)) / torch.linalg.norm(torch.tril(A0.view(N, N).t()))).item(), "bound", 10 * N * eps)
        del A0, A, Lm, R, M
    if "qr" in which:
        m, n = 65536, 4096
        A0 = torch.empty(m * n, dtype=torch.float64, device=dev); A = torch.empty_like(A0); dg = torch.empty(n, dtype=torch.float64, device=dev)
        _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), m, n, m, 8, s))
        for it in range(2):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            l0 = L.na_kernel_launches()
            e0.record(); st = L.na_qr_f64_dev(m, n, A.data_ptr(), m, dg.data_ptr(), s); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); fl = 2.0 * m * n * n - 2.0 * n ** 3 / 3
            print(f"qr {m}x{n}: status {st} {ms:.2f} ms {fl/ms/1e9:.2f} TFLOP/s ({fl/ms/1e9/37.18*100:.1f}%) launches {L.na_kernel_launches()-l0}")
        Q = torch.empty(m * n, dtype=torch.float64, device=dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _capi.check(L.na_qr_q_f64_dev(m, n, A.data_ptr(), m, dg.data_ptr(), Q.data_ptr(), m, s)); e1.record(); torch.cuda.synchronize()
        print(f"  q(): {e0.elapsed_time(e1):.2f} ms")
        Qm = Q.view(n, m).t(); R = torch.triu(A.view(n, m).t()[:n, :]); R.diagonal().copy_(dg.abs())
        A0m = A0.view(n, m).t()
        print("  resid |A-QR|/|A|", (torch.linalg.norm(A0m - Qm @ R) / torch.linalg.norm(A0m)).item(), "orth", torch.linalg.norm(Qm.t() @ Qm - torch.eye(n, device=dev, dtype=torch.float64)).item(), "bound", 10 * m * eps)
        del A0, A, Q, Qm, R
    if "lu" in which:
        A0 = torch.empty(N * N, dtype=torch.float64, device=dev)
        _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 6, s))
        A = torch.empty_like(A0); swaps = (C.c_size_t * (2 * N))(); ns = C.c_size_t(0)
        for it in range(3):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            l0 = L.na_kernel_launches()
            e0.record(); st = L.na_lu_f64_dev(N, N, A.data_ptr(), N, swaps, C.addressof(ns), s); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"lu N={N}: status {st} {ms:.2f} ms {2*N**3/3/ms/1e9:.2f} TFLOP/s ({2*N**3/3/ms/1e9/37.18*100:.1f}%) nswaps {ns.value} launches {L.na_kernel_launches()-l0}")
        LU = A.view(N, N).t()
        Lm = torch.tril(LU, -1); Lm.diagonal().add_(1.0); Um = torch.triu(LU)
        PA = A0.view(N, N).t().clone()
        sw = np.frombuffer(swaps, dtype=np.uint64)[: 2 * ns.value].reshape(-1, 2).astype(np.int64)
        perm = np.arange(N)
        for i, j in sw: perm[[i, j]] = perm[[j, i]]
        PA = PA[torch.from_numpy(perm).to(dev)]
        print("  resid |PA-LU|/|A|", (torch.linalg.norm(PA - Lm @ Um) / torch.linalg.norm(PA)).item(), "bound", 10 * N * eps)
