"""Multi-GPU factorizations: 1D column-block-cyclic layout, one process per GPU, panels broadcast
with NCCL over NVLink (BASELINE configs[2]: Cholesky N=65536 across 8 B200).

Layout.  Block column b (width nb, the last one may be narrower) lives on rank ``b % world``.  A
rank stores its block columns back to back as one column-major ``n x local_cols`` matrix (ld = n).

Cholesky step k (right-looking):
  1. owner(k): POTRF of the diagonal block (``na_cholesky_f64_dev``: nalgebra's pivot rule, failure
     column), then the panel below it  A21 <- A21 * L11^-T  (``na_trsm_f64_dev``);
  2. the (n - k*nb) x nb panel (+ a status word) is broadcast from owner(k);
  3. every rank updates its own block columns b > k with the lower-trapezoid GEMM
     ``na_dgemm_lower_dev`` (C -= P[rows >= b] * P[b]^T): the strict upper triangle is never touched,
     exactly like the reference (src/linalg/cholesky.rs:226-235).
Look-ahead: the owner of block k+1 updates that block first, factors it and issues its broadcast on
a side stream while all ranks finish the step-k updates of their other blocks.

The local arithmetic is behind a small ``ops`` object so that the ownership / broadcast logic can be
exercised on CPU with gloo (tests/test_multirank_cpu.py plugs in the oracle); the product path is
``DeviceOps`` = the C ABI of libnalgebra_b200.so on CUDA tensors.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _capi
from .sharding import block_cyclic_owner


class DeviceOps:
    """Local operations through the C ABI on device memory (torch only provides the buffers)."""

    def __init__(self, device: torch.device):
        self.device = device
        self.lib = _capi.lib()
        _capi.check(self.lib.na_init(device.index if device.index is not None else 0))

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def empty(self, numel: int) -> torch.Tensor:
        return torch.empty(numel, dtype=torch.float64, device=self.device)

    def fill_spd(self, ptr: int, nrows: int, ncols: int, ld: int, seed: int, row0: int, col0: int, n: int):
        _capi.check(self.lib.na_fill_spd_block_dev(ptr, nrows, ncols, ld, seed, row0, col0, n, self._stream()))

    def potrf(self, ptr: int, w: int, ld: int) -> int:
        fail = C.c_size_t(0)
        return _capi.check(self.lib.na_cholesky_f64_dev(w, ptr, ld, 0, 0.0, C.addressof(fail), self._stream()))

    def trsm_right_lower_trans(self, m: int, w: int, t_ptr: int, ldt: int, b_ptr: int, ldb: int):
        _capi.check(self.lib.na_trsm_f64_dev(1, 1, 1, 0, m, w, t_ptr, ldt, b_ptr, ldb, self._stream()))

    def reserve_sms(self, n_reserved: int):
        """Keep `n_reserved` SMs free of (persistent) GEMM CTAs, e.g. for NCCL's copy kernels; 0 = none."""
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        _capi.check(self.lib.na_set_gemm_sm_limit(sms - n_reserved if n_reserved > 0 else 0))

    def syrk_lower_update(self, m: int, k: int, n: int, p_ptr: int, ldp: int, c_ptr: int, ldc: int):
        """C (m x n lower trapezoid) -= P[0:m, :] * P[0:n, :]^T."""
        _capi.check(self.lib.na_dgemm_lower_dev(m, k, n, -1.0, p_ptr, 1, ldp, p_ptr, ldp, 1, 1.0, c_ptr, ldc, self._stream()))


class ColumnBlockCyclic:
    """n x n f64 matrix, column blocks of width nb dealt round-robin to the ranks."""

    def __init__(self, n: int, nb: int, rank: int, world: int, ops):
        self.n, self.nb, self.rank, self.world, self.ops = n, nb, rank, world, ops
        self.nblocks = (n + nb - 1) // nb
        self.my_blocks = [b for b in range(self.nblocks) if block_cyclic_owner(b, world) == rank]
        self.col_off, off = {}, 0
        for b in self.my_blocks:
            self.col_off[b] = off
            off += self.width(b)
        self.local_cols = off
        self.data = ops.empty(max(off, 1) * n)          # column-major n x local_cols, ld = n

    def width(self, b: int) -> int:
        return min(self.nb, self.n - b * self.nb)

    def ptr(self, row: int, b: int) -> int:
        return self.data.data_ptr() + 8 * (row + self.col_off[b] * self.n)

    def block_view(self, b: int, row0: int = 0) -> torch.Tensor:
        """(width, n - row0) torch view = the column-major (n - row0) x width sub-block."""
        lc = self.col_off[b]
        return self.data[: self.local_cols * self.n].view(self.local_cols, self.n)[lc: lc + self.width(b), row0:]

    def fill_spd(self, seed: int):
        for b in self.my_blocks:
            self.ops.fill_spd(self.ptr(0, b), self.n, self.width(b), self.n, seed, 0, b * self.nb, self.n)

    def gather_to(self, dst_rank: int = 0, group=None):
        """Full matrix on dst_rank as an (n, n) tensor in column-major element order (testing only)."""
        parts = [None] * self.world if self.rank == dst_rank else None
        dist.gather_object((self.my_blocks, self.data[: self.local_cols * self.n].cpu()), parts, dst=dst_rank, group=group)
        if self.rank != dst_rank:
            return None
        full = torch.empty(self.n, self.n, dtype=torch.float64)      # full[j, i] = A(i, j)
        for blocks, data in parts:
            data = data.view(-1, self.n)
            off = 0
            for b in blocks:
                w = self.width(b)
                full[b * self.nb: b * self.nb + w, :] = data[off: off + w, :]
                off += w
        return full.t()


def cholesky_block_cyclic(A: ColumnBlockCyclic, group=None, lookahead: bool = True) -> int:
    """In-place lower Cholesky of the distributed SPD matrix.  Returns 0 (NA_OK) or 1 (NA_NOT_PD) on
    every rank.  Only the lower triangle is read or written."""
    n, nb, rank, world, ops = A.n, A.nb, A.rank, A.world, A.ops
    bufs = [ops.empty(n * nb + 1), ops.empty(n * nb + 1)]
    use_side = lookahead and A.data.is_cuda and world > 1
    side = torch.cuda.Stream(device=A.data.device) if use_side else None
    pending = None          # (work handle or None, buffer index) of the panel broadcast in flight

    def factor_and_pack(k: int, buf: torch.Tensor) -> None:
        r0, w = k * nb, A.width(k)
        rows = n - r0
        st = ops.potrf(A.ptr(r0, k), w, n)
        if rows > w and st == 0:
            ops.trsm_right_lower_trans(rows - w, w, A.ptr(r0, k), n, A.ptr(r0 + w, k), n)
        buf[: rows * w].view(w, rows).copy_(A.block_view(k, r0))
        buf[rows * w] = float(st)

    def bcast(k: int, buf: torch.Tensor):
        r0, w = k * nb, A.width(k)
        rows = n - r0
        if world == 1:
            return None
        return dist.broadcast(buf[: rows * w + 1], src=block_cyclic_owner(k, world), group=group, async_op=True)

    def update_block(b: int, k: int, buf: torch.Tensor) -> None:
        r0, w = k * nb, A.width(k)
        rows = n - r0
        c0, wb = b * nb, A.width(b)
        ops.syrk_lower_update(n - c0, w, wb, buf.data_ptr() + 8 * (c0 - r0), rows, A.ptr(c0, b), n)

    # panel 0
    if block_cyclic_owner(0, world) == rank:
        factor_and_pack(0, bufs[0])
    pending = (bcast(0, bufs[0]), 0)
    status = 0
    for k in range(A.nblocks):
        work, bi = pending
        if work is not None:
            work.wait()
        buf = bufs[bi]
        rows_k, w_k = n - k * nb, A.width(k)
        st = int(buf[rows_k * w_k].item())
        if st != 0:
            status = 1
            break
        mine = [b for b in A.my_blocks if b > k]
        nxt = k + 1
        if nxt < A.nblocks:
            nbuf = bufs[1 - bi]
            if block_cyclic_owner(nxt, world) == rank:
                update_block(nxt, k, buf)                  # look-ahead: the next panel's block first
                factor_and_pack(nxt, nbuf)
                mine = [b for b in mine if b != nxt]
            if use_side:
                side.wait_stream(torch.cuda.current_stream(A.data.device))
                with torch.cuda.stream(side):
                    pending = (bcast(nxt, nbuf), 1 - bi)
            else:
                pending = (bcast(nxt, nbuf), 1 - bi)
        for b in mine:
            update_block(b, k, buf)
        if use_side:
            torch.cuda.current_stream(A.data.device).wait_stream(side)
    if A.data.is_cuda:
        torch.cuda.synchronize(A.data.device)
    return status
