"""GPU, >= 2 devices (-m gpu; skipped on a single-GPU box): the block-cyclic multi-GPU Cholesky through
the C ABI + NCCL matches the single-GPU factorization."""
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
    from nalgebra_b200.distributed import ColumnBlockCyclic, DeviceOps, cholesky_block_cyclic
    A = ColumnBlockCyclic(n, nb, rank, world, DeviceOps(torch.device(f"cuda:{rank}")))
    A.fill_spd(5)
    st = cholesky_block_cyclic(A)
    full = A.gather_to(0)
    if rank == 0:
        np.save(os.path.join(out_dir, "l.npy"), full.numpy())
        np.save(os.path.join(out_dir, "st.npy"), np.array([st]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_block_cyclic_cholesky_matches_single_gpu(tmp_path, nab, oracle):
    n, nb, world = 3000, 256, 2
    mp.spawn(_worker, args=(world, 29631, n, nb, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "l.npy")
    assert np.load(tmp_path / "st.npy")[0] == 0
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
    from nalgebra_b200 import _capi
    from nalgebra_b200.distributed import ColumnBlockCyclic, DeviceOps, lu_block_cyclic
    dev = torch.device(f"cuda:{rank}")
    ops = DeviceOps(dev)
    A = ColumnBlockCyclic(n, nb, rank, world, ops)
    s = torch.cuda.current_stream().cuda_stream
    for b in A.my_blocks:
        _capi.check(ops.lib.na_fill_uniform_block_dev(A.ptr(0, b), n, A.width(b), n, 6, 0, b * nb, n, s))
    pairs = lu_block_cyclic(A)
    full = A.gather_to(0)
    if rank == 0:
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
