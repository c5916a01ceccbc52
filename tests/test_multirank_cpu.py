"""CPU, world_size 2 over gloo: the N>1 host logic of the sharded GEMM -- every rank derives its
tile from (rank, world), fills its panels with the shared counter-based generator, and the tiles
assemble to the single-process product.  (The device kernels are exercised by the -m gpu tests; the
arithmetic here is the oracle's, as the checker.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from nalgebra_b200 import sharding as S
    r0, r1, c0, c1 = S.gemm_tile(rank, world, n, n)
    # this rank's panels, generated in place from global indices (what na_fill_uniform_block_dev does on the GPU)
    idx_a = (np.arange(r0, r1)[:, None] + np.arange(n)[None, :] * n).astype(np.uint64)
    idx_b = (np.arange(n)[:, None] + np.arange(c0, c1)[None, :] * n).astype(np.uint64)
    a = np.asfortranarray(O.rand01(1, idx_a.ravel()).reshape(idx_a.shape))
    b = np.asfortranarray(O.rand01(2, idx_b.ravel()).reshape(idx_b.shape))
    c = np.zeros((r1 - r0, c1 - c0), order="F")
    O.gemm(1.0, a, b, 0.0, c)
    # exchange: checksums all-reduced, tiles gathered on rank 0
    chk = torch.tensor([c.sum()], dtype=torch.float64)
    dist.all_reduce(chk)
    tiles = [None] * world
    dist.all_gather_object(tiles, (r0, r1, c0, c1, c))
    if rank == 0:
        full = np.zeros((n, n))
        for (a0, a1, b0, b1, t) in tiles:
            full[a0:a1, b0:b1] = t
        np.save(os.path.join(out_dir, "c.npy"), full)
        np.save(os.path.join(out_dir, "chk.npy"), chk.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gemm_two_ranks(tmp_path, oracle):
    n, world, port = 96, 2, 29613
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "c.npy")
    a, b = oracle.uniform(n, n, 1), oracle.uniform(n, n, 2)
    ref = np.zeros((n, n), order="F")
    oracle.gemm(1.0, a, b, 0.0, ref)
    assert np.abs(full - ref).max() <= 4 * n * np.finfo(np.float64).eps * np.linalg.norm(a) * np.linalg.norm(b)
    assert abs(np.load(tmp_path / "chk.npy")[0] - ref.sum()) <= 1e-9 * abs(ref.sum())


# ---------------------------------------------------------------------------------------------------
# block-cyclic distributed Cholesky: ownership / look-ahead / broadcast logic on CPU (gloo, world 2 and 3)
# with the oracle plugged in as the local arithmetic
# ---------------------------------------------------------------------------------------------------
class _OracleOps:
    """Same interface as nalgebra_b200.distributed.DeviceOps, CPU tensors + the oracle (checker side)."""

    def __init__(self):
        import ctypes as C
        import oracle as O
        self.C, self.O = C, O

    def _np(self, ptr, count):
        return np.frombuffer((self.C.c_double * count).from_address(ptr), dtype=np.float64)

    def _mat(self, ptr, rows, cols, ld):
        flat = self._np(ptr, ld * (cols - 1) + rows)
        return np.lib.stride_tricks.as_strided(flat, shape=(rows, cols), strides=(8, 8 * ld))

    def empty(self, numel):
        return torch.zeros(numel, dtype=torch.float64)

    def fill_spd(self, ptr, nrows, ncols, ld, seed, row0, col0, n):
        i = (np.arange(row0, row0 + nrows)[:, None]).astype(np.uint64); j = (np.arange(col0, col0 + ncols)[None, :]).astype(np.uint64)
        bij = self.O.rand01(seed, (i + j * np.uint64(n)).ravel()).reshape(nrows, ncols)
        bji = self.O.rand01(seed, (j + i * np.uint64(n)).ravel()).reshape(nrows, ncols)
        self._mat(ptr, nrows, ncols, ld)[...] = (bij + bji) * 0.5 + n * (i == j)

    def status_word(self):
        return torch.full((1,), -1, dtype=torch.int64)

    def ipiv(self, n):
        return torch.zeros(max(n, 1), dtype=torch.int32)

    def fill_uniform(self, ptr, nrows, ncols, ld, seed, row0, col0, global_rows):
        i = (np.arange(row0, row0 + nrows)[:, None]).astype(np.uint64); j = (np.arange(col0, col0 + ncols)[None, :]).astype(np.uint64)
        self._mat(ptr, nrows, ncols, ld)[...] = self.O.rand01(seed, (i + j * np.uint64(global_rows)).ravel()).reshape(nrows, ncols)

    def potrf_async(self, ptr, w, ld, fail, col_offset):
        fc = self.C.c_size_t(0)
        st = self.O.lib().na_oracle_cholesky_f64(w, ptr, ld, 0, 0.0, self.C.addressof(fc))
        if st != 0:
            cur = int(fail[0]) & ((1 << 64) - 1)
            new = min(cur, fc.value + col_offset)
            fail[0] = new if new < (1 << 63) else new - (1 << 64)

    def trsm_right_lower_trans(self, m, w, t_ptr, ldt, b_ptr, ldb):
        import scipy.linalg as sl
        t = np.tril(self._mat(t_ptr, w, w, ldt)); b = self._mat(b_ptr, m, w, ldb)
        b[...] = sl.solve_triangular(t, b.T, lower=True).T

    def lu_panel_async(self, ptr, m, w, ld, ipiv_ptr):
        mn = min(m, w)
        swaps = (self.C.c_size_t * (2 * max(mn, 1)))(); ns = self.C.c_size_t(0)
        self.O.lib().na_oracle_lu_f64(m, w, ptr, ld, swaps, self.C.addressof(ns))
        piv = np.frombuffer((self.C.c_int32 * mn).from_address(ipiv_ptr), dtype=np.int32)
        piv[...] = np.arange(mn)
        for i in range(ns.value):
            piv[swaps[2 * i]] = swaps[2 * i + 1]

    def apply_ipiv(self, ptr, nrows, ld, ncols, ipiv_ptr, k, row0):
        if ncols == 0 or k == 0:
            return
        piv = np.frombuffer((self.C.c_int32 * k).from_address(ipiv_ptr), dtype=np.int32)
        a = self._mat(ptr, nrows, ncols, ld)
        for s_ in range(k):
            i, j = row0 + s_, row0 + int(piv[s_])
            if i != j:
                a[[i, j]] = a[[j, i]]

    def gemm(self, m, k, n, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc):
        a = self._mat(a_ptr, m, k, lda); b = self._mat(b_ptr, k, n, ldb); c = self._mat(c_ptr, m, n, ldc)
        c[...] = alpha * (a @ b) + (beta * c if beta != 0.0 else 0.0)

    def trsm_left_unit_lower(self, m, n, t_ptr, ldt, b_ptr, ldb):
        import scipy.linalg as sl
        t = np.tril(self._mat(t_ptr, m, m, ldt), -1) + np.eye(m); b = self._mat(b_ptr, m, n, ldb)
        b[...] = sl.solve_triangular(t, b, lower=True, unit_diagonal=True)

    def gemm_update(self, m, k, n, a_ptr, lda, b_ptr, ldb, c_ptr, ldc):
        a = self._mat(a_ptr, m, k, lda); b = self._mat(b_ptr, k, n, ldb); c = self._mat(c_ptr, m, n, ldc)
        c -= a @ b

    def syrk_lower_update(self, m, k, n, p_ptr, ldp, c_ptr, ldc):
        p = self._mat(p_ptr, m, k, ldp); c = self._mat(c_ptr, m, n, ldc)
        upd = p @ p[:n, :].T
        mask = np.arange(m)[:, None] >= np.arange(n)[None, :]
        c[mask] -= upd[mask]


def _chol_worker(rank, world, port, n, nb, out_dir, lookahead):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nalgebra_b200.distributed import ColumnBlockCyclic, cholesky_block_cyclic
    A = ColumnBlockCyclic(n, nb, rank, world, _OracleOps())
    A.fill_spd(5)
    st = cholesky_block_cyclic(A, lookahead=lookahead)
    from nalgebra_b200.distributed import cholesky_residual_block_cyclic
    res = cholesky_residual_block_cyclic(A, 5)
    # a non-PD matrix: every rank must report the same failing column, with no host round trip per panel
    B = ColumnBlockCyclic(n, nb, rank, world, _OracleOps())
    B.fill_spd(5)
    bad = 37
    if bad // nb in B.my_blocks:
        _OracleOps()._mat(B.ptr(0, bad // nb), n, B.width(bad // nb), n)[bad, bad % nb] = -1.0
    st_bad = cholesky_block_cyclic(B, lookahead=lookahead)
    full = A.gather_to(0)
    if rank == 0:
        np.save(os.path.join(out_dir, f"l_{world}_{int(lookahead)}.npy"), full.numpy())
        np.save(os.path.join(out_dir, f"st_{world}_{int(lookahead)}.npy"), np.array([st, res, st_bad, -1 if B.fail_col is None else B.fail_col]))
    dist.barrier()
    dist.destroy_process_group()


def test_block_cyclic_cholesky_logic(tmp_path, oracle):
    n, nb = 100, 16                      # 7 blocks, the last one narrow
    spd = oracle.spd_wellcond(n, 5)
    lref = np.tril(oracle.cholesky(spd))
    for world, port in ((2, 29621), (3, 29622)):
        mp.spawn(_chol_worker, args=(world, port, n, nb, str(tmp_path), True), nprocs=world, join=True)
        got = np.load(tmp_path / f"l_{world}_1.npy")
        st, res, st_bad, fail_col = np.load(tmp_path / f"st_{world}_1.npy")
        assert st == 0 and res <= 10 * n * np.finfo(np.float64).eps
        assert st_bad == 1 and fail_col == 37                            # NA_NOT_PD + the failing column, read back once
        assert np.abs(np.tril(got) - lref).max() <= 1e-12 * np.abs(lref).max()
        assert np.array_equal(np.triu(got, 1), np.triu(spd, 1))          # strict upper never touched


def _lu_worker(rank, world, port, n, nb, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from nalgebra_b200.distributed import ColumnBlockCyclic, lu_block_cyclic
    ops = _OracleOps()
    A = ColumnBlockCyclic(n, nb, rank, world, ops)
    full0 = O.uniform(n, n, 6) - 0.3
    for b in A.my_blocks:                                   # this rank's columns of the shared test matrix
        w = A.width(b)
        ops._mat(A.ptr(0, b), n, w, n)[...] = full0[:, b * nb: b * nb + w]
    pairs = lu_block_cyclic(A)
    full = A.gather_to(0)
    # the residual replay against the generator-defined matrix (what the 8-GPU bench runs at N = 65536)
    from nalgebra_b200.distributed import lu_residual_block_cyclic
    G = ColumnBlockCyclic(n, nb, rank, world, ops)
    G.fill_uniform(6)
    lu_block_cyclic(G)
    res = lu_residual_block_cyclic(G, 6)
    if rank == 0:
        np.save(os.path.join(out_dir, f"lures_{world}.npy"), np.array([res]))
        np.save(os.path.join(out_dir, f"lu_{world}.npy"), full.numpy())
        np.save(os.path.join(out_dir, f"sw_{world}.npy"), np.array(pairs, dtype=np.int64).reshape(-1, 2))
    dist.barrier()
    dist.destroy_process_group()


def test_block_cyclic_lu_logic(tmp_path, oracle):
    n, nb = 100, 16
    a = oracle.uniform(n, n, 6) - 0.3
    lu_ref, sw_ref = oracle.lu(a)
    for world, port in ((2, 29623), (3, 29624)):
        mp.spawn(_lu_worker, args=(world, port, n, nb, str(tmp_path)), nprocs=world, join=True)
        got = np.load(tmp_path / f"lu_{world}.npy"); sw = np.load(tmp_path / f"sw_{world}.npy")
        assert np.array_equal(sw, sw_ref.astype(np.int64))                # PermutationSequence identical
        assert np.abs(got - lu_ref).max() <= 1e-11
        assert np.load(tmp_path / f"lures_{world}.npy")[0] <= 10 * n * np.finfo(np.float64).eps


# ---------------------------------------------------------------------------------------------------
# Gemm2D (the product's sharded GEMM) on CPU: collective exchange over gloo, oracle arithmetic, 2 and 4 ranks
# ---------------------------------------------------------------------------------------------------
def _gemm2d_worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nalgebra_b200.distributed import Gemm2D
    g = Gemm2D(n, n, n, rank, world, _OracleOps(), exchange="collective", kp=16)
    g.fill_uniform(1, 2)
    c = g.multiply().clone()
    c2 = g.multiply_assembled().clone()
    tiles = [None] * world
    dist.all_gather_object(tiles, (g.row0, g.col0, g.m_loc, g.n_loc, c.numpy(), float((c - c2).abs().max())))
    if rank == 0:
        full = np.zeros((n, n))
        for (r0, c0, ml, nl, t, d) in tiles:
            full[r0:r0 + ml, c0:c0 + nl] = t.reshape(nl, ml).T          # column-major tile
            assert d <= 1e-12
        np.save(os.path.join(out_dir, f"g2d_{world}.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_gemm2d_product_logic(tmp_path, oracle):
    n = 64
    a, b = oracle.uniform(n, n, 1), oracle.uniform(n, n, 2)
    ref = np.zeros((n, n), order="F")
    oracle.gemm(1.0, a, b, 0.0, ref)
    for world, port in ((2, 29631), (4, 29632)):
        mp.spawn(_gemm2d_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
        full = np.load(tmp_path / f"g2d_{world}.npy")
        assert np.abs(full - ref).max() <= 4 * n * np.finfo(np.float64).eps * np.linalg.norm(a) * np.linalg.norm(b)


def test_nccl_sm_reservation_rule(monkeypatch):
    """distributed._nccl_sm_reserve: as many SMs as NCCL has channels when the caller capped them, nothing otherwise."""
    import torch
    from nalgebra_b200 import distributed as D

    class Fake:
        def __init__(self, world, cuda):
            self.world = world
            self.data = type("T", (), {"is_cuda": cuda})()
    for k in ("NAB_BC_RESERVE", "NCCL_MAX_NCHANNELS"):
        monkeypatch.delenv(k, raising=False)
    assert D._nccl_sm_reserve(Fake(8, True)) == 0                  # NCCL's default channel count: no reservation
    monkeypatch.setenv("NCCL_MAX_NCHANNELS", "8")
    assert D._nccl_sm_reserve(Fake(8, True)) == 8
    assert D._nccl_sm_reserve(Fake(1, True)) == 0 and D._nccl_sm_reserve(Fake(8, False)) == 0
    monkeypatch.setenv("NCCL_MAX_NCHANNELS", "32")
    assert D._nccl_sm_reserve(Fake(8, True)) == 0                  # too many to give away
    monkeypatch.setenv("NAB_BC_RESERVE", "12")
    assert D._nccl_sm_reserve(Fake(8, True)) == 12
