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

    def lu_panel(self, ptr: int, m: int, w: int, ld: int):
        """LU::new on the m x w panel in place; returns the swaps as a list of (i, i2) panel-relative pairs."""
        mn = min(m, w)
        swaps = (C.c_size_t * (2 * max(mn, 1)))()
        ns = C.c_size_t(0)
        _capi.check(self.lib.na_lu_f64_dev(m, w, ptr, ld, swaps, C.addressof(ns), self._stream()))
        return [(swaps[2 * i], swaps[2 * i + 1]) for i in range(ns.value)]

    def permute_rows(self, ptr: int, nrows: int, ld: int, ncols: int, pairs):
        if not pairs or ncols == 0:
            return
        arr = (C.c_size_t * (2 * len(pairs)))(*[v for p in pairs for v in p])
        _capi.check(self.lib.na_permute_rows_f64_dev(nrows, ptr, ld, ncols, arr, len(pairs), 0, self._stream()))

    def trsm_left_unit_lower(self, m: int, n: int, t_ptr: int, ldt: int, b_ptr: int, ldb: int):
        _capi.check(self.lib.na_trsm_f64_dev(0, 1, 0, 1, m, n, t_ptr, ldt, b_ptr, ldb, self._stream()))

    def gemm_update(self, m: int, k: int, n: int, a_ptr: int, lda: int, b_ptr: int, ldb: int, c_ptr: int, ldc: int):
        """C -= A * B."""
        _capi.check(self.lib.na_dgemm_dev(m, k, n, -1.0, a_ptr, 1, lda, b_ptr, 1, ldb, 1.0, c_ptr, 1, ldc, self._stream()))

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


def lu_block_cyclic(A: ColumnBlockCyclic, group=None, lookahead: bool = True):
    """In-place LU with partial pivoting of the distributed matrix (LU::new semantics: whole rows are
    swapped, packed L\\U layout).  Returns the PermutationSequence pairs [(i, i2), ...] (global row
    indices, application order), identical on every rank.

    Step k: owner(k) factors its (n - k*nb) x nb panel locally -- in a 1D column layout the panel is
    local, so there is no cross-GPU pivot search --, broadcasts the panel and its pivot pairs; every rank
    applies the row swaps to all its other columns, solves U12 = L11^-1 A12 on its trailing columns and
    updates them with one GEMM (local block columns right of k are contiguous in local storage)."""
    n, nb, rank, world, ops = A.n, A.nb, A.rank, A.world, A.ops
    bufs = [ops.empty(n * nb), ops.empty(n * nb)]
    use_side = lookahead and A.data.is_cuda and world > 1
    side = torch.cuda.Stream(device=A.data.device) if use_side else None
    piv_cap = 2 * nb + 1
    pivs = [torch.zeros(piv_cap, dtype=torch.int64, device=A.data.device) for _ in range(2)]
    all_pairs = []

    def factor_and_pack(k: int, buf, piv):
        r0, w = k * nb, A.width(k)
        rows = n - r0
        pairs = ops.lu_panel(A.ptr(r0, k), rows, w, n)
        buf[: rows * w].view(w, rows).copy_(A.block_view(k, r0))
        host = torch.zeros(piv_cap, dtype=torch.int64)
        host[0] = len(pairs)
        for i, (a, b) in enumerate(pairs):
            host[1 + 2 * i] = a + r0
            host[2 + 2 * i] = b + r0
        piv.copy_(host)

    def bcast(k: int, buf, piv):
        if world == 1:
            return []
        r0, w = k * nb, A.width(k)
        src = block_cyclic_owner(k, world)
        return [dist.broadcast(buf[: (n - r0) * w], src=src, group=group, async_op=True),
                dist.broadcast(piv, src=src, group=group, async_op=True)]

    def local_range(b_lo: int, b_hi: int):
        """(local column offset, number of columns) of the local blocks b with b_lo <= b < b_hi."""
        blocks = [b for b in A.my_blocks if b_lo <= b < b_hi]
        if not blocks:
            return 0, 0
        return A.col_off[blocks[0]], sum(A.width(b) for b in blocks)

    def apply_panel(k: int, buf, pairs, b_lo: int, b_hi: int, swap_only: bool):
        """Row swaps (+ TRSM + GEMM unless swap_only) of panel k on the local blocks in [b_lo, b_hi)."""
        off, ncols = local_range(b_lo, b_hi)
        if ncols == 0:
            return
        r0, w = k * nb, A.width(k)
        rows = n - r0
        base = A.data.data_ptr() + 8 * off * n
        ops.permute_rows(base, n, n, ncols, pairs)
        if swap_only:
            return
        ops.trsm_left_unit_lower(w, ncols, buf.data_ptr(), rows, base + 8 * r0, n)
        if rows > w:
            ops.gemm_update(rows - w, w, ncols, buf.data_ptr() + 8 * w, rows, base + 8 * r0, n, base + 8 * (r0 + w), n)

    if block_cyclic_owner(0, world) == rank:
        factor_and_pack(0, bufs[0], pivs[0])
    pending = (bcast(0, bufs[0], pivs[0]), 0)
    for k in range(A.nblocks):
        works, bi = pending
        for wk in works:
            wk.wait()
        buf, piv = bufs[bi], pivs[bi]
        ph = piv.cpu()
        pairs = [(int(ph[1 + 2 * i]), int(ph[2 + 2 * i])) for i in range(int(ph[0]))]
        all_pairs.extend(pairs)
        # the owner's panel block is already swapped; swap the local blocks left of k
        apply_panel(k, buf, pairs, 0, k, True)
        nxt = k + 1
        lo = k + 1
        if nxt < A.nblocks:
            nbi = 1 - bi
            if block_cyclic_owner(nxt, world) == rank:
                apply_panel(k, buf, pairs, nxt, nxt + 1, False)      # look-ahead: the next panel's block first
                factor_and_pack(nxt, bufs[nbi], pivs[nbi])
                lo = nxt + 1
            if use_side:
                side.wait_stream(torch.cuda.current_stream(A.data.device))
                with torch.cuda.stream(side):
                    pending = (bcast(nxt, bufs[nbi], pivs[nbi]), nbi)
            else:
                pending = (bcast(nxt, bufs[nbi], pivs[nbi]), nbi)
        apply_panel(k, buf, pairs, lo, A.nblocks, False)
        if use_side:
            torch.cuda.current_stream(A.data.device).wait_stream(side)
    if A.data.is_cuda:
        torch.cuda.synchronize(A.data.device)
    return all_pairs
