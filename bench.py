#!/usr/bin/env python
"""bench.py -- headline benchmark of nalgebra_b200 (contract: see the task prompt / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): f64 GEMM GFLOP/s at N=16384 (C = A*B, DMatrix<f64> 16384x16384x16384),
reported with the fraction of the measured FP64 peak; Cholesky / LU / QR numbers of the same run
ride along under "extra".  One "step" = one full 16384^3 GEMM.

* value      : device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
* e2e        : the same GEMM through the host-pointer C-ABI call `na_dgemm` (what nalgebra's
               gemm_uninit would call instead of matrixmultiply::dgemm) with pinned HOST buffers;
               H2D of A and B and D2H of C are inside the timed region.
* roofline   : DMMA tensor-pipe roofline; peak = 37.18 TFLOP/s measured on this pool's B200 by
               tools/fp64_peak.cu (profiles/fp64_peak_r01.md) -- MEASURED_PEAKS.json carries no FP64
               number, so this is "of measured (own DMMA micro-benchmark)".
* cpu_baseline / --impl reference : the oracle's matrixmultiply restatement on the host cores
               (kind "port": the Rust reference cannot be built in this image).
* N > 1      : 2D output-tile sharding of the same 16384^3 problem over an r x c process grid
               (strong scaling), one process per GPU under torchrun.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL = 16384
FP64_PEAK_TFLOPS = 37.18           # measured, profiles/fp64_peak_r01.md (DMMA m8n8k4, 148 SMs @ 1965 MHz)
GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v == "Active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(x for x in sm if x > 0.5 * mx) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's matrixmultiply restatement (test infrastructure used as the baseline)
# ------------------------------------------------------------------------------------------------
def cpu_gemm_gflops(n: int, nthreads: int, reps: int = 1):
    import numpy as np
    import oracle as O
    a = O.uniform(n, n, 1); b = O.uniform(n, n, 2); c = np.zeros((n, n), order="F")
    O.gemm(1.0, a[:256, :256].copy(order="F"), b[:256, :256].copy(order="F"), 0.0, c[:256, :256].copy(order="F"), path="mm")  # warm the lib
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        O.gemm(1.0, a, b, 0.0, c, path="mm", nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return 2.0 * n ** 3 / best / 1e9, best


def run_reference(args):
    """--impl reference: nalgebra's CPU path for this metric.  The Rust reference cannot be built
    here (no cargo/rustc), so this times the oracle port of gemm_uninit -> matrixmultiply::dgemm on
    a bounded sample of the workload, with all host threads (the crate's optional `threading`
    feature; nalgebra's default is 1 thread, see cpu_baseline in the main arm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 4096
    times = []
    for i in range(args.warmup + args.steps):
        gf, dt = cpu_gemm_gflops(n, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = 2.0 * n ** 3 / (ms * 1e-3) / 1e9
    sample = f"{n}^3 f64 GEMM per step (1/64 of the 16384^3 workload), oracle matrixmultiply port, {cores} threads"
    line = {
        "impl": "reference", "metric": "f64_gemm_gflops_n16384", "value": value, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"DMatrix<f64> GEMM {N_FULL}x{N_FULL}x{N_FULL} (C = A*B), uniform [0,1) inputs",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from nalgebra_b200 import _capi
    from nalgebra_b200.sharding import gemm_piece_class, gemm_piece_order, process_grid

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if world != args.gpus and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    ngpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    L = _capi.lib()
    _capi.check(L.na_init(local_rank))
    stream = torch.cuda.current_stream().cuda_stream
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count

    N = args.n
    pr, pc = process_grid(ngpus)
    my_r, my_c = rank // pc, rank % pc
    m_loc, n_loc = N // pr, N // pc
    row0, col0 = my_r * m_loc, my_c * n_loc

    # Inputs resident in HBM, block-distributed WITHOUT replication:
    #   rank (r, c) owns the K-chunk c of the row panel A[r-rows, :]   (m_loc x N/pc, contiguous)
    #   and the K-chunk r of the column panel B[:, c-cols]             (N/pr x n_loc, contiguous).
    # A step assembles the panels over NVLink -- ONE in-place all-gather per panel, in the row / column
    # sub-communicator, both issued up front -- and accumulates C over K-pieces as they become available:
    # fully local pieces first, then those that only need the (smaller, earlier) B gather, then the rest.
    # The exchange overlaps the DMMA work; while a gather is outstanding the persistent GEMM leaves 20 SMs
    # to NCCL's kernels.  (N = 1: no exchange.)  Per-chunk broadcasts (4 + 2 collectives, the later ones
    # queued behind full-grid GEMM pieces) cost 7.4 ms per step at 8 GPUs.
    kca, kcb = N // pc, N // pr
    A = torch.empty(m_loc * N, dtype=torch.float64, device=dev)                       # full row panel, ld = m_loc
    Bfull = torch.empty(pr * kcb * n_loc, dtype=torch.float64, device=dev)
    Bc = [Bfull[q * kcb * n_loc:(q + 1) * kcb * n_loc] for q in range(pr)]              # K-chunks of the column panel
    Cd = torch.empty(m_loc * n_loc, dtype=torch.float64, device=dev)
    a_chunks = [A[q * kca * m_loc:(q + 1) * kca * m_loc] for q in range(pc)]
    _capi.check(L.na_fill_uniform_block_dev(a_chunks[my_c].data_ptr(), m_loc, kca, m_loc, 1, row0, my_c * kca, N, stream))
    _capi.check(L.na_fill_uniform_block_dev(Bc[my_r].data_ptr(), kcb, n_loc, kcb, 2, my_r * kcb, col0, N, stream))
    row_group = col_group = None
    if world > 1:
        for r in range(pr):
            g = dist.new_group([r * pc + c for c in range(pc)])
            if r == my_r:
                row_group = g
        for c in range(pc):
            g = dist.new_group([r * pc + c for r in range(pr)])
            if c == my_c:
                col_group = g
    kp = min(kca, kcb, 2048 if world > 1 else N)   # K-piece of one GEMM call (small pieces: only the first runs on a reduced grid)
    piece_ms = 2.0 * m_loc * n_loc * kp / 36e12 * 1e3
    n_limited = max(1, int(-(-5.0 // piece_ms)))          # pieces that start within ~5 ms of the step
    def piece_class(t):      # 0: both operands local, 1: only B foreign, 2: only A foreign, 3: both foreign
        return gemm_piece_class(t, kp, kca, kcb, my_r, my_c)
    pieces = gemm_piece_order(N, kp, kca, kcb, my_r, my_c)

    def gemm_piece(t, first):
        k0 = t * kp
        qa, qb = k0 // kca, k0 // kcb
        a_ptr = a_chunks[qa].data_ptr() + 8 * (k0 - qa * kca) * m_loc
        b_ptr = Bc[qb].data_ptr() + 8 * (k0 - qb * kcb)
        _capi.check(L.na_dgemm_dev(m_loc, kp, n_loc, 1.0, a_ptr, 1, m_loc, b_ptr, 1, kcb, 0.0 if first else 1.0,
                                   Cd.data_ptr(), 1, m_loc, stream))

    def step():
        ha = hb = None
        if world > 1:
            if pr > 1:
                hb = dist.all_gather_into_tensor(Bfull, Bc[my_r], group=col_group, async_op=True)
            if pc > 1:
                ha = dist.all_gather_into_tensor(A, a_chunks[my_c], group=row_group, async_op=True)
        for i, t in enumerate(pieces):
            cls = piece_class(t)
            if hb is not None and cls in (1, 3):
                hb.wait(); hb = None
            if ha is not None and cls in (2, 3):
                ha.wait(); ha = None
            # while a gather can still be in flight (the first ~5 ms of the step: 0.27 - 1.07 GB over NVLink), leave SMs
            # to NCCL's kernels: the GEMM is persistent, a full grid would make the CTAs that find no SM start late
            L.na_set_gemm_sm_limit(sm_count - 20 if (i < n_limited and (ha is not None or hb is not None)) else 0)
            gemm_piece(t, i == 0)
        if ha is not None:
            ha.wait()
        if hb is not None:
            hb.wait()
        L.na_set_gemm_sm_limit(0)

    def step_replicated():
        """Panels already assembled (what `step` leaves in A / Bc): compute only, one call per B chunk."""
        for q in range(pr):
            _capi.check(L.na_dgemm_dev(m_loc, kcb, n_loc, 1.0, A.data_ptr() + 8 * q * kcb * m_loc, 1, m_loc,
                                       Bc[q].data_ptr(), 1, kcb, 0.0 if q == 0 else 1.0, Cd.data_ptr(), 1, m_loc, stream))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # the sampler starts before the warm-up (nvidia-smi needs ~0.2 s to produce its first line); only
    # samples taken under load (SM clock above half of max) enter the summary
    with ClockSampler(local_rank) as clk:
        for _ in range(args.warmup):
            step()
        barrier()
        if ngpus > 1 and args.steps * 0.25 / ngpus < 0.5:      # short timed region: keep the GPU busy long enough to be sampled
            for _ in range(max(3, int(0.6 / (0.25 / ngpus)))):
                step()
            barrier()
        launches0 = L.na_kernel_launches()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    launches = L.na_kernel_launches() - launches0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    flops = 2.0 * N ** 3
    value = flops / (ms_step * 1e-3) / 1e9
    clocks = clk.summary()

    # roofline of the dominant kernel: compute-only launches (panels assembled), CUDA events on the launching stream
    nk = len(pieces) if ngpus > 1 else 1
    if ngpus > 1:
        step_replicated(); torch.cuda.synchronize()
        k0e = torch.cuda.Event(enable_timing=True); k1e = torch.cuda.Event(enable_timing=True)
        k0e.record()
        for _ in range(args.steps):
            step_replicated()
        k1e.record(); torch.cuda.synchronize()
        kernel_ms = k0e.elapsed_time(k1e) / args.steps
        nk = pr
    else:
        kernel_ms = ms_total / args.steps
    exchange_check = None
    if ngpus > 1:      # the exchanged + K-pieced product against the product of the assembled panels (same kernel, other K order)
        step(); torch.cuda.synchronize()
        c_step = Cd.clone()
        step_replicated(); torch.cuda.synchronize()
        exchange_check = float((c_step - Cd).abs().max().item() / max(Cd.abs().max().item(), 1e-300))
        del c_step
    achieved = (flops / ngpus) / (kernel_ms * 1e-3) / 1e12
    traffic = None
    try:   # DRAM bytes of one launch from the committed `ncu --set full` capture (same workload only)
        with open(os.path.join(ROOT, "profiles", "r01_dgemm_traffic.json")) as f:
            tj = json.load(f)
        if ngpus == 1 and N == N_FULL:
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass

    # ---- e2e: host-pointer C-ABI call, pinned host buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hA = torch.empty(m_loc * N, dtype=torch.float64).pin_memory()
        hB = torch.empty(N * n_loc, dtype=torch.float64).pin_memory()
        hC = torch.empty(m_loc * n_loc, dtype=torch.float64).pin_memory()
        hA.copy_(A)                                                     # assembled row panel (ld = m_loc)
        hBv = hB.view(n_loc, N)                                         # column-major N x n_loc
        for q in range(pr):
            hBv[:, q * kcb:(q + 1) * kcb].copy_(Bc[q].view(n_loc, kcb))
        torch.cuda.synchronize()

        def e2e_step():
            _capi.check(L.na_dgemm(m_loc, N, n_loc, 1.0, hA.data_ptr(), 1, m_loc, hB.data_ptr(), 1, N, 0.0,
                                   hC.data_ptr(), 1, m_loc))

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()                      # blocks until C is visible on the host
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = tt.item() / e2e_steps * 1e3
        checksum = float(hC[:1024].sum())
        e2e = {"value": flops / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int((hA.numel() + hB.numel()) * 8), "d2h_bytes_per_step": int(hC.numel() * 8),
               "steps": e2e_steps, "api": "na_dgemm (host pointers, pinned)", "checksum": checksum}
        del hA, hB, hC

    extra = {}
    if ngpus == 1 and not args.no_extra:
        try:
            extra = factorization_extras(L, _capi, torch, dev, stream, N)
        except Exception as ex:  # a missing entry point must not kill the headline line
            extra = {"error": repr(ex)}

    if ngpus > 1 and not args.no_extra:
        # BASELINE configs[2]: Cholesky, 1D block-cyclic over the ranks, panels broadcast with NCCL
        try:
            from nalgebra_b200.distributed import ColumnBlockCyclic, DeviceOps, cholesky_block_cyclic
            n_bc = {2: 32768, 4: 49152, 8: 65536}.get(ngpus, 16384)
            nb_bc = 1024
            Abc = ColumnBlockCyclic(n_bc, nb_bc, rank, world, DeviceOps(dev))
            best = None
            for _ in range(2):
                Abc.fill_spd(5)
                barrier()
                t0 = time.perf_counter()
                st_bc = cholesky_block_cyclic(Abc)
                torch.cuda.synchronize()
                tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                best = tt.item() if best is None else min(best, tt.item())
            gf = (n_bc ** 3 / 3.0) / best / 1e9
            extra["cholesky_block_cyclic"] = {"n": n_bc, "nb": nb_bc, "ms": best * 1e3, "gflops": gf, "status": st_bc,
                                              "pct_of_fp64_peak": gf / 1e3 / (FP64_PEAK_TFLOPS * ngpus) * 100,
                                              "exchange": "NCCL broadcast of each factored panel from its owner"}
            # BASELINE configs[3] at scale: LU with partial pivoting, same layout, panel + pivot pairs broadcast
            from nalgebra_b200.distributed import lu_block_cyclic
            best = None
            for _ in range(2):
                for b in Abc.my_blocks:
                    _capi.check(L.na_fill_uniform_block_dev(Abc.ptr(0, b), n_bc, Abc.width(b), n_bc, 6, 0, b * nb_bc, n_bc, stream))
                barrier()
                t0 = time.perf_counter()
                pairs = lu_block_cyclic(Abc)
                torch.cuda.synchronize()
                tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                best = tt.item() if best is None else min(best, tt.item())
            gf = (2.0 * n_bc ** 3 / 3.0) / best / 1e9
            extra["lu_block_cyclic"] = {"n": n_bc, "nb": nb_bc, "ms": best * 1e3, "gflops": gf, "nswaps": len(pairs),
                                        "pct_of_fp64_peak": gf / 1e3 / (FP64_PEAK_TFLOPS * ngpus) * 100,
                                        "exchange": "NCCL broadcast of each factored panel and its pivot pairs from the owner"}
            del Abc
        except Exception as ex:
            extra.setdefault("cholesky_block_cyclic", {"error": repr(ex)})
            extra.setdefault("lu_block_cyclic", {"error": repr(ex)})

    cpu = None
    if rank == 0 and ngpus == 1 and not args.no_cpu:
        n_s = 4096
        gf, dt = cpu_gemm_gflops(n_s, 1)
        cpu = {"value": gf, "unit": "GFLOP/s", "cores": 1, "kind": "port",
               "sample": f"{n_s}^3 f64 GEMM (1/64 of the workload), oracle matrixmultiply port, 1 thread = nalgebra's default features; {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": "f64_gemm_gflops_n16384", "value": value, "unit": "GFLOP/s", "n_gpus": ngpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"DMatrix<f64> GEMM {N}x{N}x{N} (C = A*B), uniform [0,1) inputs (BASELINE.json configs[1])",
                       "process_grid": f"{pr}x{pc}", "per_gpu_tile": f"{m_loc}x{n_loc}x{N}",
                       "distribution": ("single GPU" if ngpus == 1 else
                                        f"A and B block-distributed without replication; per step each rank receives {pc - 1} A K-chunks "
                                        f"({(pc - 1) * kca * m_loc * 8 / 1e9:.2f} GB) and {pr - 1} B K-chunks ({(pr - 1) * kcb * n_loc * 8 / 1e9:.2f} GB) "
                                        "over NVLink (one in-place NCCL all-gather per panel in the row / column sub-communicator) overlapped with the K-chunked GEMM"),
                       "l2": "inputs (>=1.6 GB per GPU) exceed the 126 MB L2; no explicit flush",
                       "pct_of_fp64_peak": 100.0 * value / 1e3 / (FP64_PEAK_TFLOPS * ngpus)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": achieved / FP64_PEAK_TFLOPS, "traffic": traffic,
                         "kernel": "dgemm_tma_dmma_kernel", "peak_source": "measured: tools/fp64_peak.cu DMMA chain on this pool's B200 (profiles/fp64_peak_r01.md); MEASURED_PEAKS.json has no FP64 entry",
                         "algorithmic_flops_per_launch": flops / ngpus / nk, "launches_per_step": nk,
                         "compute_only_ms_per_step": kernel_ms},
            "clocks": clocks, "gpu_launches": int(launches),
        }
        if exchange_check is not None:
            line["config"]["exchange_check_max_rel_diff"] = exchange_check     # exchanged K-pieced product vs assembled panels
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def factorization_extras(L, _capi, torch, dev, stream, N):
    """Cholesky / LU (/ QR) at the BASELINE sizes on one GPU, device resident, CUDA-event timed."""
    import ctypes as C
    out = {}

    if hasattr(L, "na_cholesky_f64_dev"):
        A = torch.empty(N * N, dtype=torch.float64, device=dev)
        A0 = torch.empty(N * N, dtype=torch.float64, device=dev)
        _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 5, stream))
        M = A0.view(N, N)
        M.copy_((M + M.t()) * 0.5); M.diagonal().add_(float(N))       # (B+B^T)/2 + n*I, SURVEY §8(d) Cfg 3 (ii)
        fail = C.c_size_t(0)
        best = None
        for _ in range(3):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            st = _capi.check(L.na_cholesky_f64_dev(N, A.data_ptr(), N, 0, 0.0, C.addressof(fail), stream))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
        gf = (N ** 3 / 3.0) / (best * 1e-3) / 1e9
        out["cholesky_n16384"] = {"ms": best, "gflops": gf, "pct_of_fp64_peak": gf / 1e3 / FP64_PEAK_TFLOPS * 100, "status": st}
        del A, A0
    if hasattr(L, "na_lu_f64_dev"):
        A = torch.empty(N * N, dtype=torch.float64, device=dev)
        A0 = torch.empty(N * N, dtype=torch.float64, device=dev)
        _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 6, stream))
        swaps = (C.c_size_t * (2 * N))(); ns = C.c_size_t(0)
        best = None
        for _ in range(3):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            _capi.check(L.na_lu_f64_dev(N, N, A.data_ptr(), N, swaps, C.addressof(ns), stream))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
        gf = (2.0 * N ** 3 / 3.0) / (best * 1e-3) / 1e9
        out["lu_n16384"] = {"ms": best, "gflops": gf, "pct_of_fp64_peak": gf / 1e3 / FP64_PEAK_TFLOPS * 100, "nswaps": ns.value}
        del A, A0
    if hasattr(L, "na_qr_f64_dev"):
        m, n = 65536, 4096
        A = torch.empty(m * n, dtype=torch.float64, device=dev)
        A0 = torch.empty(m * n, dtype=torch.float64, device=dev)
        diag = torch.empty(n, dtype=torch.float64, device=dev)
        _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), m, n, m, 8, stream))
        best = None
        for _ in range(2):
            A.copy_(A0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            _capi.check(L.na_qr_f64_dev(m, n, A.data_ptr(), m, diag.data_ptr(), stream))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
        fl = 2.0 * m * n * n - 2.0 * n ** 3 / 3.0
        gf = fl / (best * 1e-3) / 1e9
        out["qr_65536x4096"] = {"ms": best, "gflops": gf, "pct_of_fp64_peak": gf / 1e3 / FP64_PEAK_TFLOPS * 100}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_FULL, help="problem size (default: the BASELINE config, 16384)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
