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
