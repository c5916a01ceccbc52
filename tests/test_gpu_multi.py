"""GPU, >= 2 devices (-m gpu; skipped on a single-GPU box): the multi-GPU paths through the C ABI + NCCL / peer copies --
block-cyclic Cholesky and LU against the single-GPU factorization and the oracle (pivots bit-exact), their residual
replays, and the 2D-sharded GEMM (both exchange back ends) against the oracle.
A log of a 2-GPU run of this file is kept under profiles/ (the driver's test box has one GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, nb, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from nalgebra_b200.distributed import ColumnBlockCyclic, DeviceOps, cholesky_block_cyclic, cholesky_residual_block_cyclic
    ops = DeviceOps(torch.device(f"cuda:{rank}"))
    A = ColumnBlockCyclic(n, nb, rank, world, ops)
    A.fill_spd(5)
    st = cholesky_block_cyclic(A)
    res = cholesky_residual_block_cyclic(A, 5)
    # not positive definite: one negative diagonal entry -> NA_NOT_PD + the failing column on every rank
    B = ColumnBlockCyclic(n, nb, rank, world, ops)
    B.fill_spd(5)
    bad = 1234
    if bad // nb in B.my_blocks:
        B.block_view(bad // nb)[bad % nb, bad] = -1.0
    st_bad = cholesky_block_cyclic(B)
    full = A.gather_to(0)
    if rank == 0:
        np.save(os.path.join(out_dir, "l.npy"), full.numpy())
        np.save(os.path.join(out_dir, "st.npy"), np.array([st, res, st_bad, -1 if B.fail_col is None else B.fail_col]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_block_cyclic_cholesky_matches_single_gpu(tmp_path, nab, oracle):
    n, nb, world = 3000, 256, 2
    mp.spawn(_worker, args=(world, 29631, n, nb, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "l.npy")
    st, res, st_bad, fail_col = np.load(tmp_path / "st.npy")
    assert st == 0 and res <= 10 * n * np.finfo(np.float64).eps
    assert st_bad == 1 and fail_col == 1234
    spd = oracle.spd_wellcond(n, 5)
    ch = nab.Cholesky.new(spd)
    l1 = ch.l()
    assert np.abs(np.tril(got) - l1).max() <= 1e-10 * np.abs(l1).max()
    assert np.array_equal(np.triu(got, 1), np.triu(spd, 1))
    assert np.linalg.norm(np.tril(got) @ np.tril(got).T - spd) / np.linalg.norm(spd) <= 10 * n * np.finfo(np.float64).eps


def _lu_worker(rank, world, port, n, nb, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from nalgebra_b200.distributed import ColumnBlockCyclic, DeviceOps, lu_block_cyclic, lu_residual_block_cyclic
    dev = torch.device(f"cuda:{rank}")
    ops = DeviceOps(dev)
    A = ColumnBlockCyclic(n, nb, rank, world, ops)
    A.fill_uniform(6)
    pairs = lu_block_cyclic(A)
    res = lu_residual_block_cyclic(A, 6)
    full = A.gather_to(0)
    if rank == 0:
        np.save(os.path.join(out_dir, "lures.npy"), np.array([res]))
        np.save(os.path.join(out_dir, "lu.npy"), full.numpy())
        np.save(os.path.join(out_dir, "sw.npy"), np.array(pairs, dtype=np.int64).reshape(-1, 2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_block_cyclic_lu_matches_oracle_pivots(tmp_path, nab, oracle):
    n, nb, world = 2000, 256, 2
    mp.spawn(_lu_worker, args=(world, 29632, n, nb, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "lu.npy"); sw = np.load(tmp_path / "sw.npy")
    a = oracle.uniform(n, n, 6)
    lu_ref, sw_ref = oracle.lu(a)
    assert np.array_equal(sw, sw_ref.astype(np.int64))            # pivots bit-exact with the CPU oracle
    assert np.abs(got - lu_ref).max() <= 1e-9
    assert np.load(tmp_path / "lures.npy")[0] <= 10 * n * np.finfo(np.float64).eps


def _gemm_worker(rank, world, port, n, exchange, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from nalgebra_b200.distributed import DeviceOps, Gemm2D
    g = Gemm2D(n, n, n, rank, world, DeviceOps(torch.device(f"cuda:{rank}")), exchange=exchange, kp=256)
    g.fill_uniform(1, 2)
    c1 = g.multiply().clone()
    c2 = g.multiply().clone()                       # a second step on the same (peer-mapped) buffers
    tiles = [None] * world
    dist.all_gather_object(tiles, (g.row0, g.col0, g.m_loc, g.n_loc, c1.cpu().numpy(), bool(torch.equal(c1, c2))))
    if rank == 0:
        full = np.zeros((n, n))
        for (r0, c0, ml, nl, t, same) in tiles:
            assert same
            full[r0:r0 + ml, c0:c0 + nl] = t.reshape(nl, ml).T
        np.save(os.path.join(out_dir, f"c_{exchange}.npy"), full)
    g.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange,port", [("p2p", 29641), ("collective", 29642)])
def test_gemm2d_vs_oracle(tmp_path, nab, oracle, exchange, port):
    """The product's sharded GEMM (nalgebra_b200.distributed.Gemm2D): tiles assembled from 2 GPUs == the oracle's product
    within the north-star bound, for the copy-engine peer exchange and for the NCCL all-gather exchange."""
    n, world = 1024, 2
    mp.spawn(_gemm_worker, args=(world, port, n, exchange, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / f"c_{exchange}.npy")
    a, b = oracle.uniform(n, n, 1), oracle.uniform(n, n, 2)
    ref = np.zeros((n, n), order="F")
    oracle.gemm(1.0, a, b, 0.0, ref, path="mm", nthreads=8)
    assert np.abs(got - ref).max() <= 4 * n * np.finfo(np.float64).eps * np.linalg.norm(a) * np.linalg.norm(b)
