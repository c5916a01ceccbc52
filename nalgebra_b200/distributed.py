"""Multi-GPU paths of the hot path, one process per GPU under torchrun.

* ``Gemm2D``: C = A * B sharded as 2D output tiles over a pr x pc process grid (BASELINE configs[1]).  A and
  B are block-distributed without replication; each ``multiply()`` stages the missing K-chunks of the row /
  column panel over NVLink while the K-pieces whose operands are already local run on the DMMA engine.
  Exchange back ends: ``"p2p"`` -- peer copies by the copy engines out of CUDA-IPC mapped buffers (no SM is
  taken from the persistent GEMM), ``"collective"`` -- one in-place all-gather per panel (NCCL on GPUs, gloo in
  the CPU logic tests).
* ``cholesky_block_cyclic`` / ``lu_block_cyclic``: 1D column-block-cyclic factorizations (configs[2], [3] at
  scale); panels (and LU's pivot vector) broadcast from their owner.  Nothing on the per-panel path touches the
  host: the Cholesky status word and LU's pivots stay on the device (``na_cholesky_f64_dev_async``,
  ``na_lu_f64_dev_async``, ``na_apply_ipiv_f64_dev``) and are read back once at the end.
* ``cholesky_residual_block_cyclic`` / ``lu_residual_block_cyclic``: ||A - L L^T||_F / ||A||_F and
  ||P A - L U||_F / ||A||_F evaluated on the distributed layout by replaying the panel broadcasts against a
  regenerated A (the north-star correctness gate at sizes no single GPU or CPU oracle can hold).

Layout of the factorizations.  Block column b (width nb, the last one may be narrower) lives on rank
``b % world``.  A rank stores its block columns back to back as one column-major ``n x local_cols`` matrix.

The local arithmetic is behind a small ``ops`` object so that the ownership / broadcast logic can be exercised
on CPU with gloo (tests/test_multirank_cpu.py plugs in the oracle); the product path is ``DeviceOps`` = the C
ABI of libnalgebra_b200.so on CUDA memory.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import os

import torch
import torch.distributed as dist

from . import _capi
from .sharding import block_cyclic_owner, gemm_piece_class, gemm_piece_order, process_grid

U64_MAX = (1 << 64) - 1


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

    def status_word(self) -> torch.Tensor:
        """Device word for the failing column of the (async) Cholesky panels: all ones = no failure."""
        return torch.full((1,), -1, dtype=torch.int64, device=self.device)

    def ipiv(self, n: int) -> torch.Tensor:
        return torch.zeros(max(n, 1), dtype=torch.int32, device=self.device)

    def fill_spd(self, ptr: int, nrows: int, ncols: int, ld: int, seed: int, row0: int, col0: int, n: int):
        _capi.check(self.lib.na_fill_spd_block_dev(ptr, nrows, ncols, ld, seed, row0, col0, n, self._stream()))

    def fill_uniform(self, ptr: int, nrows: int, ncols: int, ld: int, seed: int, row0: int, col0: int, global_rows: int):
        _capi.check(self.lib.na_fill_uniform_block_dev(ptr, nrows, ncols, ld, seed, row0, col0, global_rows, self._stream()))

    def potrf_async(self, ptr: int, w: int, ld: int, fail: torch.Tensor, col_offset: int):
        _capi.check(self.lib.na_cholesky_f64_dev_async(w, ptr, ld, 0, 0.0, fail.data_ptr(), col_offset, self._stream()))

    def trsm_right_lower_trans(self, m: int, w: int, t_ptr: int, ldt: int, b_ptr: int, ldb: int):
        _capi.check(self.lib.na_trsm_f64_dev(1, 1, 1, 0, m, w, t_ptr, ldt, b_ptr, ldb, self._stream()))

    def reserve_sms(self, n_reserved: int):
        """Keep `n_reserved` SMs free of (persistent) GEMM CTAs, e.g. for NCCL's copy kernels; 0 = none."""
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        _capi.check(self.lib.na_set_gemm_sm_limit(sms - n_reserved if n_reserved > 0 else 0))

    def lu_panel_async(self, ptr: int, m: int, w: int, ld: int, ipiv_ptr: int):
        """LU::new on the m x w panel in place; the 0-based panel-relative pivot rows go to the device ints at ipiv_ptr."""
        _capi.check(self.lib.na_lu_f64_dev_async(m, w, ptr, ld, ipiv_ptr, self._stream()))

    def apply_ipiv(self, ptr: int, nrows: int, ld: int, ncols: int, ipiv_ptr: int, k: int, row0: int):
        if ncols == 0 or k == 0:
            return
        _capi.check(self.lib.na_apply_ipiv_f64_dev(nrows, ptr, ld, ncols, ipiv_ptr, k, row0, self._stream()))

    def trsm_left_unit_lower(self, m: int, n: int, t_ptr: int, ldt: int, b_ptr: int, ldb: int):
        _capi.check(self.lib.na_trsm_f64_dev(0, 1, 0, 1, m, n, t_ptr, ldt, b_ptr, ldb, self._stream()))

    def gemm(self, m: int, k: int, n: int, alpha: float, a_ptr: int, lda: int, b_ptr: int, ldb: int, beta: float, c_ptr: int, ldc: int):
        """C <- alpha * A * B + beta * C (column-major operands)."""
        _capi.check(self.lib.na_dgemm_dev(m, k, n, alpha, a_ptr, 1, lda, b_ptr, 1, ldb, beta, c_ptr, 1, ldc, self._stream()))

    def gemm_update(self, m: int, k: int, n: int, a_ptr: int, lda: int, b_ptr: int, ldb: int, c_ptr: int, ldc: int):
        """C -= A * B."""
        self.gemm(m, k, n, -1.0, a_ptr, lda, b_ptr, ldb, 1.0, c_ptr, ldc)

    def syrk_lower_update(self, m: int, k: int, n: int, p_ptr: int, ldp: int, c_ptr: int, ldc: int):
        """C (m x n lower trapezoid) -= P[0:m, :] * P[0:n, :]^T."""
        _capi.check(self.lib.na_dgemm_lower_dev(m, k, n, -1.0, p_ptr, 1, ldp, p_ptr, ldp, 1, 1.0, c_ptr, ldc, self._stream()))


# ---------------------------------------------------------------------------------------------------
# GEMM: 2D output tiles
# ---------------------------------------------------------------------------------------------------
class _RawDeviceBuffer:
    """cudaMalloc'ed memory (CUDA-IPC exportable, unlike a caching-allocator block) seen by torch through
    __cuda_array_interface__."""

    def __init__(self, lib, numel: int):
        self.lib, self.numel = lib, numel
        p = C.c_void_p()
        _capi.check(lib.na_dev_malloc(C.byref(p), max(numel, 1) * 8))
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": (max(numel, 1),), "typestr": "<f8", "data": (self.ptr, False), "version": 2}

    def tensor(self, device) -> torch.Tensor:
        return torch.as_tensor(self, device=device)

    def free(self):
        if self.ptr:
            self.lib.na_dev_free(self.ptr)
            self.ptr = None


class Gemm2D:
    """C (m x n) = A (m x k) * B (k x n), 2D output-tile sharding over a pr x pc grid (``process_grid(world)``).

    Rank (r, c) owns the tile C[r-rows, c-cols] and, WITHOUT replication, K-chunk c of the row panel A[r-rows, :]
    (m_loc x k/pc) and K-chunk r of the column panel B[:, c-cols] (k/pr x n_loc).  ``multiply()`` assembles the two
    panels from the ranks of the same grid row / column and accumulates the tile over K-pieces in the order
    local / only-B-foreign / only-A-foreign / both (``sharding.gemm_piece_order``), so the exchange overlaps the
    DMMA work.  Reference seam: one `matrixmultiply::dgemm` call per tile (src/base/blas_uninit.rs:298-313)."""

    def __init__(self, m: int, n: int, k: int, rank: int, world: int, ops, *, exchange: str = "auto", kp: int = 2048, group=None):
        self.m, self.n, self.k, self.rank, self.world, self.ops = m, n, k, rank, world, ops
        self.pr, self.pc = process_grid(world)
        if m % self.pr or n % self.pc or k % self.pr or k % self.pc:
            raise ValueError("Gemm2D: m, n, k must be divisible by the process grid")
        self.my_r, self.my_c = rank // self.pc, rank % self.pc
        self.m_loc, self.n_loc = m // self.pr, n // self.pc
        self.row0, self.col0 = self.my_r * self.m_loc, self.my_c * self.n_loc
        self.kca, self.kcb = k // self.pc, k // self.pr
        self.kp = min(self.kca, self.kcb, kp if world > 1 else k)
        while self.kca % self.kp or self.kcb % self.kp:
            self.kp //= 2
        self.pieces = gemm_piece_order(k, self.kp, self.kca, self.kcb, self.my_r, self.my_c)
        self.device = getattr(ops, "device", torch.device("cpu"))
        self.on_gpu = self.device.type == "cuda"
        if exchange == "auto":
            exchange = "p2p" if (self.on_gpu and world > 1) else "collective"
        self.exchange = exchange if world > 1 else "none"
        self.group = group
        self.row_group = self.col_group = None
        self._raw = []
        if self.exchange == "p2p":
            lib = ops.lib
            ra, rb = _RawDeviceBuffer(lib, self.m_loc * k), _RawDeviceBuffer(lib, k * self.n_loc)
            self._raw = [ra, rb]
            self.A, self.Bfull = ra.tensor(self.device), rb.tensor(self.device)
        else:
            self.A = ops.empty(self.m_loc * k)               # full row panel, column-major m_loc x k (ld = m_loc)
            self.Bfull = ops.empty(k * self.n_loc)           # K-chunks of the column panel, each kcb x n_loc (ld = kcb)
        self.C = ops.empty(self.m_loc * self.n_loc)
        self.a_chunks = [self.A[q * self.kca * self.m_loc:(q + 1) * self.kca * self.m_loc] for q in range(self.pc)]
        self.b_chunks = [self.Bfull[q * self.kcb * self.n_loc:(q + 1) * self.kcb * self.n_loc] for q in range(self.pr)]
        if world > 1:
            for r in range(self.pr):
                g = dist.new_group([r * self.pc + c for c in range(self.pc)])
                if r == self.my_r:
                    self.row_group = g
            for c in range(self.pc):
                g = dist.new_group([r * self.pc + c for r in range(self.pr)])
                if c == self.my_c:
                    self.col_group = g
        if self.exchange == "p2p":
            self._open_peers()

    # -- inputs ---------------------------------------------------------------------------------
    def own_a(self) -> torch.Tensor:
        """This rank's K-chunk of the row panel: m_loc x kca, column-major, global origin (row0, my_c * kca)."""
        return self.a_chunks[self.my_c]

    def own_b(self) -> torch.Tensor:
        """This rank's K-chunk of the column panel: kcb x n_loc, column-major, global origin (my_r * kcb, col0)."""
        return self.b_chunks[self.my_r]

    def fill_uniform(self, seed_a: int, seed_b: int):
        """A(i, j) = rand01(seed_a, i + j*m), B(i, j) = rand01(seed_b, i + j*k): the generator shared with the oracle."""
        self.ops.fill_uniform(self.own_a().data_ptr(), self.m_loc, self.kca, self.m_loc, seed_a, self.row0, self.my_c * self.kca, self.m)
        self.ops.fill_uniform(self.own_b().data_ptr(), self.kcb, self.n_loc, self.kcb, seed_b, self.my_r * self.kcb, self.col0, self.k)

    # -- peer mapping (exchange == "p2p") -----------------------------------------------------------
    def _open_peers(self):
        lib = self.ops.lib
        handles = []
        for raw in self._raw:
            h = (C.c_ubyte * 64)()
            _capi.check(lib.na_ipc_get_handle(raw.ptr, h))
            handles.append(bytes(h))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=self.group)
        self.peer_a, self.peer_b = {}, {}
        for c in range(self.pc):                       # the ranks of my grid row hold the other K-chunks of my A panel
            if c != self.my_c:
                p = C.c_void_p()
                _capi.check(lib.na_ipc_open_handle((C.c_ubyte * 64).from_buffer_copy(everyone[self.my_r * self.pc + c][0]), C.byref(p)))
                self.peer_a[c] = p.value
        for r in range(self.pr):                       # the ranks of my grid column hold the other K-chunks of my B panel
            if r != self.my_r:
                p = C.c_void_p()
                _capi.check(lib.na_ipc_open_handle((C.c_ubyte * 64).from_buffer_copy(everyone[r * self.pc + self.my_c][1]), C.byref(p)))
                self.peer_b[r] = p.value
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)

    def close(self):
        if self.exchange == "p2p":
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            for p in list(self.peer_a.values()) + list(self.peer_b.values()):
                self.ops.lib.na_ipc_close_handle(p)
            self.peer_a, self.peer_b = {}, {}
            dist.barrier(group=self.group)
        del self.A, self.Bfull, self.a_chunks, self.b_chunks
        for raw in self._raw:
            raw.free()
        self._raw = []

    # -- the product ------------------------------------------------------------------------------
    def _piece(self, t: int, first: bool, alpha: float, beta: float):
        k0 = t * self.kp
        qa, qb = k0 // self.kca, k0 // self.kcb
        a_ptr = self.a_chunks[qa].data_ptr() + 8 * (k0 - qa * self.kca) * self.m_loc
        b_ptr = self.b_chunks[qb].data_ptr() + 8 * (k0 - qb * self.kcb)
        self.ops.gemm(self.m_loc, self.kp, self.n_loc, alpha, a_ptr, self.m_loc, b_ptr, self.kcb, beta if first else 1.0,
                      self.C.data_ptr(), self.m_loc)

    def multiply(self, alpha: float = 1.0, beta: float = 0.0) -> torch.Tensor:
        """C_tile <- alpha * A[r-rows, :] * B[:, c-cols] + beta * C_tile.  The other ranks' K-chunks must be complete in
        their stream order when they call multiply(); on return (stream order) this rank's tile is complete and no peer
        reads this rank's inputs any more."""
        if self.exchange == "none":
            for i, t in enumerate(self.pieces):
                self._piece(t, i == 0, alpha, beta)
            return self.C
        if self.exchange == "p2p":
            return self._multiply_p2p(alpha, beta)
        return self._multiply_collective(alpha, beta)

    def _multiply_collective(self, alpha, beta):
        ha = hb = None
        if self.pr > 1:
            hb = dist.all_gather_into_tensor(self.Bfull, self.own_b(), group=self.col_group, async_op=True)
        if self.pc > 1:
            ha = dist.all_gather_into_tensor(self.A, self.own_a(), group=self.row_group, async_op=True)
        # While a gather can still be in flight (the first ~5 ms: up to 1.07 GB over NVLink) the persistent GEMM leaves
        # SMs to NCCL's kernels -- a full grid would make the CTAs that find no SM start late.
        piece_ms = 2.0 * self.m_loc * self.n_loc * self.kp / 36e12 * 1e3
        n_limited = max(1, int(-(-5.0 // piece_ms)))
        for i, t in enumerate(self.pieces):
            cls = gemm_piece_class(t, self.kp, self.kca, self.kcb, self.my_r, self.my_c)
            if hb is not None and cls in (1, 3):
                hb.wait(); hb = None
            if ha is not None and cls in (2, 3):
                ha.wait(); ha = None
            if self.on_gpu:
                self.ops.reserve_sms(20 if (i < n_limited and (ha is not None or hb is not None)) else 0)
            self._piece(t, i == 0, alpha, beta)
        for h in (ha, hb):
            if h is not None:
                h.wait()
        if self.on_gpu:
            self.ops.reserve_sms(0)
        return self.C

    def _multiply_p2p(self, alpha, beta):
        cur = torch.cuda.current_stream(self.device)
        lib = self.ops.lib
        # every rank's inputs are complete (stream order) once this tiny all-reduce has run
        dist.all_reduce(self._flag, group=self.group)
        self.copy_stream.wait_stream(cur)
        ev_b, ev_a = torch.cuda.Event(), torch.cuda.Event()
        cs = self.copy_stream.cuda_stream
        # peer copies by the copy engines: B first (smaller; needed first by the piece order), then A
        nb_bytes, na_bytes = self.kcb * self.n_loc * 8, self.kca * self.m_loc * 8
        for r, base in self.peer_b.items():
            _capi.check(lib.na_memcpy_peer_async(self.b_chunks[r].data_ptr(), base + r * nb_bytes, nb_bytes, cs))
        ev_b.record(self.copy_stream)
        for c, base in self.peer_a.items():
            _capi.check(lib.na_memcpy_peer_async(self.a_chunks[c].data_ptr(), base + c * na_bytes, na_bytes, cs))
        ev_a.record(self.copy_stream)
        got_a = got_b = False
        for i, t in enumerate(self.pieces):
            cls = gemm_piece_class(t, self.kp, self.kca, self.kcb, self.my_r, self.my_c)
            if cls in (1, 3) and not got_b:
                cur.wait_event(ev_b); got_b = True
            if cls in (2, 3) and not got_a:
                cur.wait_event(ev_a); got_a = True
            self._piece(t, i == 0, alpha, beta)
        cur.wait_event(ev_a); cur.wait_event(ev_b)
        # nobody reads this rank's inputs after this second all-reduce (stream order)
        dist.all_reduce(self._flag, group=self.group)
        return self.C

    def multiply_assembled(self, alpha: float = 1.0, beta: float = 0.0) -> torch.Tensor:
        """Compute only, on panels that a previous multiply() left assembled: one GEMM call per B chunk."""
        for q in range(self.pr):
            self.ops.gemm(self.m_loc, self.kcb, self.n_loc, alpha, self.A.data_ptr() + 8 * q * self.kcb * self.m_loc, self.m_loc,
                          self.b_chunks[q].data_ptr(), self.kcb, beta if q == 0 else 1.0, self.C.data_ptr(), self.m_loc)
        return self.C


# ---------------------------------------------------------------------------------------------------
# 1D column-block-cyclic factorizations
# ---------------------------------------------------------------------------------------------------
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

    def local_range(self, b_lo: int, b_hi: int):
        """(local column offset, number of columns) of the local blocks b with b_lo <= b < b_hi (contiguous in storage)."""
        blocks = [b for b in self.my_blocks if b_lo <= b < b_hi]
        if not blocks:
            return 0, 0
        return self.col_off[blocks[0]], sum(self.width(b) for b in blocks)

    def fill_spd(self, seed: int):
        for b in self.my_blocks:
            self.ops.fill_spd(self.ptr(0, b), self.n, self.width(b), self.n, seed, 0, b * self.nb, self.n)

    def fill_uniform(self, seed: int):
        for b in self.my_blocks:
            self.ops.fill_uniform(self.ptr(0, b), self.n, self.width(b), self.n, seed, 0, b * self.nb, self.n)

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


class _PanelPipe:
    """Double-buffered panel broadcast with one-step look-ahead: the owner of panel k+1 produces it while all ranks
    are still applying panel k, and its broadcast travels on a side stream (GPU) meanwhile."""

    def __init__(self, A: ColumnBlockCyclic, group, lookahead: bool, extra=None):
        self.A, self.group = A, group
        n, nb = A.n, A.nb
        self.bufs = [A.ops.empty(n * nb), A.ops.empty(n * nb)]
        self.extra = extra                                   # optional per-panel companion tensor slices to broadcast
        self.use_side = lookahead and A.data.is_cuda and A.world > 1
        self.side = torch.cuda.Stream(device=A.data.device) if self.use_side else None

    def pack(self, k: int, bi: int):
        A = self.A
        r0, w = k * A.nb, A.width(k)
        rows = A.n - r0
        self.bufs[bi][: rows * w].view(w, rows).copy_(A.block_view(k, r0))

    def bcast(self, k: int, bi: int, companions=()):
        A = self.A
        if A.world == 1:
            return []
        r0, w = k * A.nb, A.width(k)
        src = block_cyclic_owner(k, A.world)
        tensors = [self.bufs[bi][: (A.n - r0) * w]] + list(companions)
        if self.use_side:
            self.side.wait_stream(torch.cuda.current_stream(A.data.device))
            with torch.cuda.stream(self.side):
                return [dist.broadcast(t, src=src, group=self.group, async_op=True) for t in tensors]
        return [dist.broadcast(t, src=src, group=self.group, async_op=True) for t in tensors]

    def join(self):
        if self.use_side:
            torch.cuda.current_stream(self.A.data.device).wait_stream(self.side)


def _nccl_sm_reserve(A: "ColumnBlockCyclic") -> int:
    """SMs kept free of the persistent GEMM CTAs while a block-cyclic factorization runs on several GPUs.  The trailing
    updates are back-to-back 148-CTA launches with a static tile schedule; NCCL's broadcast of the NEXT panel (one CTA per
    channel, resident for the whole transfer) otherwise takes SMs from them and every GEMM launched meanwhile waits for its
    displaced CTAs.  Measured on 8 B200s, Cholesky N = 65536 (`profiles/r02_bc_chol_8gpu.txt`): NCCL default channels / no
    reservation 542 ms; 2 channels / 2 SMs 601; 4 / 4 519; 8 / 8 492; 12 / 12 489; 16 / 16 510.  So: as many SMs as NCCL has
    channels when the caller capped them (NCCL_MAX_NCHANNELS -- bench.py sets 8), nothing otherwise (with NCCL's default
    channel count the reservation costs more than the stalls).  NAB_BC_RESERVE overrides."""
    if A.world == 1 or not A.data.is_cuda:
        return 0
    if "NAB_BC_RESERVE" in os.environ:
        return int(os.environ["NAB_BC_RESERVE"])
    ch = int(os.environ.get("NCCL_MAX_NCHANNELS", "0") or 0)
    return ch if 0 < ch <= 16 else 0


def cholesky_block_cyclic(A: ColumnBlockCyclic, group=None, lookahead: bool = True) -> int:
    """In-place lower Cholesky of the distributed SPD matrix.  Returns 0 (NA_OK) or 1 (NA_NOT_PD) on every rank;
    ``A.fail_col`` then holds the first failing global column (the reference returns None there,
    src/linalg/cholesky.rs:243-268; what the factor holds after a failure is unspecified, as there).  Only the lower
    triangle is read or written.

    Step k: owner(k) factors the diagonal block (nalgebra's pivot rule) and solves the panel below it
    (A21 <- A21 * L11^-T); the (n - k*nb) x nb panel is broadcast; every rank updates its block columns b > k with
    the lower-trapezoid GEMM (C -= P[rows >= b] * P[b]^T, src/linalg/cholesky.rs:226-235 restructured).  The status
    stays on the device until the end: no host synchronisation per panel."""
    n, nb, rank, world, ops = A.n, A.nb, A.rank, A.world, A.ops
    pipe = _PanelPipe(A, group, lookahead)
    fail = ops.status_word()
    reserve = _nccl_sm_reserve(A)
    if reserve:
        ops.reserve_sms(reserve)

    def factor_and_pack(k: int, bi: int) -> None:
        r0, w = k * nb, A.width(k)
        rows = n - r0
        ops.potrf_async(A.ptr(r0, k), w, n, fail, r0)
        if rows > w:
            ops.trsm_right_lower_trans(rows - w, w, A.ptr(r0, k), n, A.ptr(r0 + w, k), n)
        pipe.pack(k, bi)

    def update_block(b: int, k: int, buf: torch.Tensor) -> None:
        r0, w = k * nb, A.width(k)
        rows = n - r0
        c0, wb = b * nb, A.width(b)
        ops.syrk_lower_update(n - c0, w, wb, buf.data_ptr() + 8 * (c0 - r0), rows, A.ptr(c0, b), n)

    if block_cyclic_owner(0, world) == rank:
        factor_and_pack(0, 0)
    pending = (pipe.bcast(0, 0), 0)
    for k in range(A.nblocks):
        works, bi = pending
        for wk in works:
            wk.wait()
        buf = pipe.bufs[bi]
        mine = [b for b in A.my_blocks if b > k]
        nxt = k + 1
        if nxt < A.nblocks:
            if block_cyclic_owner(nxt, world) == rank:
                update_block(nxt, k, buf)                  # look-ahead: the next panel's block first
                factor_and_pack(nxt, 1 - bi)
                mine = [b for b in mine if b != nxt]
            pending = (pipe.bcast(nxt, 1 - bi), 1 - bi)
        for b in mine:
            update_block(b, k, buf)
        pipe.join()
    if reserve:
        ops.reserve_sms(0)
    # the word is an unsigned "first failing column" (all ones = none): combine across ranks on the host, once
    local = int(fail.cpu().item()) & U64_MAX
    if world > 1:
        everyone = [None] * world
        dist.all_gather_object(everyone, local, group=group)
        local = min(everyone)
    if A.data.is_cuda:
        torch.cuda.synchronize(A.data.device)
    A.fail_col = None if local == U64_MAX else local
    return 0 if local == U64_MAX else 1


def lu_block_cyclic(A: ColumnBlockCyclic, group=None, lookahead: bool = True):
    """In-place LU with partial pivoting of the distributed matrix (LU::new semantics: whole rows are swapped,
    packed L\\U layout).  Returns the PermutationSequence pairs [(i, i2), ...] (global row indices, application
    order), identical on every rank.

    Step k: owner(k) factors its (n - k*nb) x nb panel locally -- in a 1D column layout the panel is local, so there
    is no cross-GPU pivot search --, broadcasts the panel and its pivot vector; every rank applies the row
    interchanges to all its other columns (straight from the device-resident pivot vector), solves
    U12 = L11^-1 A12 on its trailing columns and updates them with one GEMM (local block columns right of k are
    contiguous in local storage).  The pivots are read back once, at the end."""
    n, nb, rank, world, ops = A.n, A.nb, A.rank, A.world, A.ops
    pipe = _PanelPipe(A, group, lookahead)
    ipiv = ops.ipiv(n)                                   # panel-relative 0-based pivot rows of all columns
    reserve = _nccl_sm_reserve(A)
    if reserve:
        ops.reserve_sms(reserve)

    def piv_slice(k: int) -> torch.Tensor:
        r0 = k * nb
        return ipiv[r0: r0 + min(A.width(k), n - r0)]

    def factor_and_pack(k: int, bi: int):
        r0, w = k * nb, A.width(k)
        ops.lu_panel_async(A.ptr(r0, k), n - r0, w, n, piv_slice(k).data_ptr())
        pipe.pack(k, bi)

    def apply_panel(k: int, buf, b_lo: int, b_hi: int, swap_only: bool):
        """Row swaps (+ TRSM + GEMM unless swap_only) of panel k on the local blocks in [b_lo, b_hi)."""
        off, ncols = A.local_range(b_lo, b_hi)
        if ncols == 0:
            return
        r0, w = k * nb, A.width(k)
        rows = n - r0
        base = A.data.data_ptr() + 8 * off * n
        ops.apply_ipiv(base, n, n, ncols, piv_slice(k).data_ptr(), min(w, rows), r0)
        if swap_only:
            return
        ops.trsm_left_unit_lower(w, ncols, buf.data_ptr(), rows, base + 8 * r0, n)
        if rows > w:
            ops.gemm_update(rows - w, w, ncols, buf.data_ptr() + 8 * w, rows, base + 8 * r0, n, base + 8 * (r0 + w), n)

    if block_cyclic_owner(0, world) == rank:
        factor_and_pack(0, 0)
    pending = (pipe.bcast(0, 0, [piv_slice(0)]), 0)
    for k in range(A.nblocks):
        works, bi = pending
        for wk in works:
            wk.wait()
        buf = pipe.bufs[bi]
        # the owner's panel block is already swapped; swap the local blocks left of k
        apply_panel(k, buf, 0, k, True)
        nxt = k + 1
        lo = k + 1
        if nxt < A.nblocks:
            if block_cyclic_owner(nxt, world) == rank:
                apply_panel(k, buf, nxt, nxt + 1, False)      # look-ahead: the next panel's block first
                factor_and_pack(nxt, 1 - bi)
                lo = nxt + 1
            pending = (pipe.bcast(nxt, 1 - bi, [piv_slice(nxt)]), 1 - bi)
        apply_panel(k, buf, lo, A.nblocks, False)
        pipe.join()
    if A.data.is_cuda:
        torch.cuda.synchronize(A.data.device)
    A.ipiv = ipiv
    if reserve:
        ops.reserve_sms(0)
    return pivot_pairs(ipiv.cpu().tolist(), n, nb)


def pivot_pairs(ipiv_rel, n: int, nb: int):
    """PermutationSequence pairs (i, i2), i != i2, from the panel-relative pivot vector of the block-cyclic LU
    (src/linalg/permutation_sequence.rs:84-93)."""
    pairs = []
    for i in range(n):
        i2 = (i // nb) * nb + int(ipiv_rel[i])
        if i2 != i:
            pairs.append((i, i2))
    return pairs


# ---------------------------------------------------------------------------------------------------
# residuals on the distributed layout (the correctness gate at N = 65536)
# ---------------------------------------------------------------------------------------------------
def _sumsq_lower(A: ColumnBlockCyclic) -> torch.Tensor:
    """Sum of squares of the lower triangle (incl. diagonal) held by this rank."""
    s = torch.zeros((), dtype=torch.float64, device=A.data.device)
    for b in A.my_blocks:
        c0 = b * A.nb
        v = A.block_view(b, c0)                              # (w, n - c0): [j, i] = entry (c0 + i, c0 + j)
        s += torch.triu(v).pow(2).sum()                      # i >= j
    return s


def cholesky_residual_block_cyclic(L: ColumnBlockCyclic, seed: int, group=None) -> float:
    """||A - L L^T||_F / ||A||_F over the lower triangle, A = the SPD test matrix of ``fill_spd(seed)``, L = the
    distributed factor.  Replays the panel broadcasts: R <- A; for every panel k: R_b -= L_k[rows >= b] * L_k[b]^T for
    the local blocks b >= k (the factorization's own update, diagonal block included)."""
    n, nb, rank, world, ops = L.n, L.nb, L.rank, L.world, L.ops
    R = ColumnBlockCyclic(n, nb, rank, world, ops)
    R.fill_spd(seed)
    norm_a = _sumsq_lower(R)
    pipe = _PanelPipe(L, group, False)
    for k in range(L.nblocks):
        r0, w = k * nb, L.width(k)
        rows = n - r0
        if block_cyclic_owner(k, world) == rank:
            pipe.pack(k, 0)
        for wk in pipe.bcast(k, 0):
            wk.wait()
        buf = pipe.bufs[0]
        top = buf[: rows * w].view(w, rows)[:, :w]           # [j, i] = L(r0 + i, r0 + j): the strict upper triangle of the
        top.copy_(torch.triu(top))                           # diagonal block is the caller's old data, not part of L
        for b in (b for b in R.my_blocks if b >= k):
            c0, wb = b * nb, R.width(b)
            ops.syrk_lower_update(n - c0, w, wb, buf.data_ptr() + 8 * (c0 - r0), rows, R.ptr(c0, b), n)
    t = torch.stack([_sumsq_lower(R), norm_a])
    if world > 1:
        dist.all_reduce(t, group=group)
    t = t.cpu()
    return float((t[0] / t[1]).sqrt())


def lu_residual_block_cyclic(LU: ColumnBlockCyclic, seed: int, group=None) -> float:
    """||P A - L U||_F / ||A||_F, A = the uniform test matrix of ``fill_uniform(seed)``, (L\\U, ipiv) = the result of
    ``lu_block_cyclic`` (``LU.ipiv``).  R <- P A (the recorded interchanges applied to a regenerated A); for every panel
    k (broadcast again): R[r0:, b] -= Lk * U[k-rows, b] for the local blocks b > k and R[r0:, k] -= Lk * triu(U_kk)
    on the owner, Lk = the panel's unit-lower trapezoid."""
    n, nb, rank, world, ops = LU.n, LU.nb, LU.rank, LU.world, LU.ops
    R = ColumnBlockCyclic(n, nb, rank, world, ops)
    R.fill_uniform(seed)
    norm_a = R.data[: R.local_cols * n].pow(2).sum()
    base = R.data.data_ptr()
    for k in range(LU.nblocks):
        r0 = k * nb
        kk = min(LU.width(k), n - r0)
        ops.apply_ipiv(base, n, n, R.local_cols, LU.ipiv[r0: r0 + kk].data_ptr(), kk, r0)
    pipe = _PanelPipe(LU, group, False)
    lk = ops.empty(n * nb)
    for k in range(LU.nblocks):
        r0, w = k * nb, LU.width(k)
        rows = n - r0
        if block_cyclic_owner(k, world) == rank:
            pipe.pack(k, 0)
        for wk in pipe.bcast(k, 0):
            wk.wait()
        panel = pipe.bufs[0][: rows * w].view(w, rows)       # [j, i] = panel entry (i, j)
        lkv = lk[: rows * w].view(w, rows)
        lkv.copy_(panel)
        top = lkv[:, :w]
        top.copy_(torch.triu(top, 1))                        # strictly lower in matrix coordinates
        top.diagonal().fill_(1.0)
        off, ncols = LU.local_range(k + 1, LU.nblocks)
        if ncols:
            ops.gemm_update(rows, w, ncols, lk.data_ptr(), rows, LU.data.data_ptr() + 8 * (r0 + off * n), n,
                            R.data.data_ptr() + 8 * (r0 + off * n), n)
        if block_cyclic_owner(k, world) == rank:
            ukk = ops.empty(w * w)
            ukk.view(w, w).copy_(torch.tril(panel[:, :w]))   # upper incl. diagonal in matrix coordinates
            ops.gemm_update(rows, w, w, lk.data_ptr(), rows, ukk.data_ptr(), w, R.ptr(r0, k), n)
    t = torch.stack([R.data[: R.local_cols * n].pow(2).sum(), norm_a])
    if world > 1:
        dist.all_reduce(t, group=group)
    t = t.cpu()
    return float((t[0] / t[1]).sqrt())
