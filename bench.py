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
        "config": {"workload": f"DMatrix<f64> GEMM {N_FULL}x{N_FULL}x{N_FULL} (C = A*B), uniform [0,1) inputs -- CPU arm timed on a "
                               f"{n}x{n}x{n} sample of it (rate comparison, not the same problem size)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def kernel_source_hash():
    """sha256 of the dominant kernel's source file: stamps ncu-derived numbers so that they are dropped when it changes."""
    import hashlib
    with open(os.path.join(ROOT, "nalgebra_b200", "csrc", "dgemm.cu"), "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def gemm_oracle_spot_check(torch, g2d, N, step, nrows=6, ncols=6):
    """Rows/columns of this rank's C tile against the oracle (na_oracle_dgemm_mm) on generator-defined inputs.
    Returns the largest |C_gpu - C_oracle| divided by the north-star bound 4*k*eps*max|A|*max|B|-style row/column
    norms (||a_i|| * ||b_j||); <= 1 passes."""
    import numpy as np
    import oracle as O
    step(); torch.cuda.synchronize()
    m_loc, n_loc = g2d.m_loc, g2d.n_loc
    C = g2d.C.view(n_loc, m_loc)                                # [j, i] = C(row0 + i, col0 + j)
    ri = np.unique(np.linspace(0, m_loc - 1, nrows).astype(np.int64)); cj = np.unique(np.linspace(0, n_loc - 1, ncols).astype(np.int64))
    k = np.arange(N, dtype=np.uint64)
    a_rows = np.stack([O.rand01(1, np.uint64(g2d.row0 + i) + k * np.uint64(N)) for i in ri])          # A(i, k) = rand01(1, i + k*N)
    b_cols = np.stack([O.rand01(2, k + np.uint64(g2d.col0 + j) * np.uint64(N)) for j in cj], axis=1)  # B(k, j) = rand01(2, k + j*N)
    ref = np.zeros((len(ri), len(cj)), order="F")
    O.gemm(1.0, np.asfortranarray(a_rows), np.asfortranarray(b_cols), 0.0, ref, path="mm")
    got = C[torch.as_tensor(cj)][:, torch.as_tensor(ri)].cpu().numpy().T
    bound = 4 * N * np.finfo(np.float64).eps * np.linalg.norm(a_rows, axis=1)[:, None] * np.linalg.norm(b_cols, axis=0)[None, :]
    return {"entries": int(ref.size), "max_abs_err": float(np.abs(got - ref).max()), "max_err_over_bound": float((np.abs(got - ref) / bound).max()),
            "bound": "4*k*eps*||a_i||*||b_j|| per entry (north_star GEMM gate), oracle = matrixmultiply restatement"}


def oracle_factor_baselines(which):
    """the oracle's (1-thread, unblocked, reference operation order) Cholesky / LU / QR timed at small n
    and extrapolated with the n^3 law to the benchmark sizes.  `which`: {"cholesky": [n...], "lu": [...], "qr": [...]}."""
    import numpy as np
    import oracle as O
    out = {}
    for name, sizes in which.items():
        pts = []
        for n in sizes:
            if name == "cholesky":
                a = O.spd_wellcond(n, 5); fl = n ** 3 / 3.0
                t0 = time.perf_counter(); O.cholesky(a); dt = time.perf_counter() - t0
            elif name == "lu":
                a = O.uniform(n, n, 6); fl = 2.0 * n ** 3 / 3.0
                t0 = time.perf_counter(); O.lu(a); dt = time.perf_counter() - t0
            elif name == "hessenberg":
                a = O.uniform(n, n, 6); fl = 10.0 * n ** 3 / 3.0
                t0 = time.perf_counter(); O.hessenberg(a); dt = time.perf_counter() - t0
            elif name == "symmetric_tridiagonal":
                a = O.spd_wellcond(n, 5); fl = 4.0 * n ** 3 / 3.0
                t0 = time.perf_counter(); O.symmetric_tridiagonal(a); dt = time.perf_counter() - t0
            elif name == "bidiagonal":
                a = O.uniform(n, n, 6); fl = 8.0 * n ** 3 / 3.0
                t0 = time.perf_counter(); O.bidiagonal(a); dt = time.perf_counter() - t0
            else:
                a = O.uniform(n, n, 8); fl = 4.0 * n ** 3 / 3.0
                t0 = time.perf_counter(); O.qr(a); dt = time.perf_counter() - t0
            pts.append({"n": n, "s": dt, "gflops": fl / dt / 1e9})
        rate = pts[-1]["gflops"]                            # the largest timed size carries the fit (flops / s is flat in n)
        out[name] = {"value": rate, "unit": "GFLOP/s", "cores": 1, "kind": "port", "points": pts,
                     "sample": f"oracle (line-faithful C restatement of nalgebra's unblocked {name}), 1 thread, n = {sizes}; "
                               f"n^3 law: the full-size run would take flops / {rate:.2f} GFLOP/s"}
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from nalgebra_b200 import _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL here carries only the panel broadcasts of the block-cyclic factorizations (the GEMM exchange goes through the
        # copy engines): 8 channels = 8 resident CTAs, for which nalgebra_b200.distributed keeps 8 SMs free of the persistent
        # GEMM CTAs (8-GPU Cholesky N = 65536: 542 -> 492 ms, profiles/r02_bc_chol_8gpu.txt)
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "8")
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
    # The sharded GEMM is the product's (nalgebra_b200.distributed.Gemm2D): A and B block-distributed WITHOUT replication,
    # rank (r, c) owns K-chunk c of the row panel A[r-rows, :] and K-chunk r of the column panel B[:, c-cols]; a step stages
    # the missing K-chunks over NVLink (copy-engine peer copies out of CUDA-IPC mapped buffers, or NCCL all-gathers with
    # NAB_GEMM_EXCHANGE=collective) while the K-pieces whose operands are local already run.  (N = 1: no exchange.)
    from nalgebra_b200.distributed import DeviceOps, Gemm2D
    ops = DeviceOps(dev)
    g2d = Gemm2D(N, N, N, rank, world, ops, exchange=os.environ.get("NAB_GEMM_EXCHANGE", "auto"))
    pr, pc, my_r, my_c = g2d.pr, g2d.pc, g2d.my_r, g2d.my_c
    m_loc, n_loc, row0, col0, kca, kcb = g2d.m_loc, g2d.n_loc, g2d.row0, g2d.col0, g2d.kca, g2d.kcb
    g2d.fill_uniform(1, 2)
    A, Bc, Cd, pieces = g2d.A, g2d.b_chunks, g2d.C, g2d.pieces

    def step():
        g2d.multiply()

    def step_replicated():
        g2d.multiply_assembled()

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
    # Parity of the product against the ORACLE (checker only): a few rows and columns of this rank's C tile are recomputed
    # on the host with the oracle's matrixmultiply restatement from the generator-defined inputs.
    oracle_check = None
    if not args.no_cpu:
        try:
            oracle_check = gemm_oracle_spot_check(torch, g2d, N, step)
            if world > 1:
                oc = torch.tensor([oracle_check["max_err_over_bound"]], dtype=torch.float64, device=dev)
                dist.all_reduce(oc, op=dist.ReduceOp.MAX)
                oracle_check["max_err_over_bound"] = oc.item()
            oracle_check["ok"] = oracle_check["max_err_over_bound"] <= 1.0
        except Exception as ex:
            oracle_check = {"error": repr(ex)}
    achieved = (flops / ngpus) / (kernel_ms * 1e-3) / 1e12
    traffic = None
    traffic_note = None
    try:   # DRAM bytes of one launch from the committed `ncu --set full` capture -- valid only for the kernel source it was taken with
        with open(os.path.join(ROOT, "profiles", "dgemm_traffic.json")) as f:
            tj = json.load(f)
        if ngpus == 1 and N == N_FULL:
            if tj.get("kernel_source_sha256") == kernel_source_hash():
                traffic = tj["traffic_bytes_per_launch"]
            else:
                traffic_note = "committed ncu capture is from another version of dgemm.cu: dropped"
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

        e2e_steps = max(1, min(args.steps, 10))
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
            extra = factorization_extras(L, _capi, torch, dev, stream, N, cpu=not args.no_cpu, e2e=not args.no_e2e)
        except Exception as ex:  # a missing entry point must not kill the headline line
            extra = {"error": repr(ex)}

    if ngpus > 1 and not args.no_extra:
        # BASELINE configs[2] / [3] at scale: 1D block-cyclic Cholesky and LU, panels broadcast with NCCL; every timed
        # factorization is followed by its residual on the distributed layout (north-star gate 10*n*eps), and the
        # block-cyclic LU must reproduce the single-GPU pivot sequence at the largest size one GPU holds comfortably.
        eps = 2.220446049250313e-16
        try:
            from nalgebra_b200.distributed import (ColumnBlockCyclic, cholesky_block_cyclic, cholesky_residual_block_cyclic,
                                                   lu_block_cyclic, lu_residual_block_cyclic)
            n_bc = {2: 32768, 4: 49152, 8: 65536}.get(ngpus, 16384)
            # block width: 1024 on 2 GPUs; on more ranks 512 (shorter owner chain per step, finer balance): 8 GPUs N = 65536
            # 492 ms at 512, 519 at 768, 537 at 1024 (profiles/r02_bc_chol_8gpu.txt)
            nb_bc = 1024 if ngpus <= 2 else 512

            def timed(fn):
                barrier()
                t0 = time.perf_counter()
                r = fn()
                torch.cuda.synchronize()
                tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                return r, tt.item()

            Abc = ColumnBlockCyclic(n_bc, nb_bc, rank, world, ops)
            best = None
            for _ in range(2):
                Abc.fill_spd(5)
                st_bc, dt = timed(lambda: cholesky_block_cyclic(Abc))
                best = dt if best is None else min(best, dt)
            res = cholesky_residual_block_cyclic(Abc, 5)
            gf = (n_bc ** 3 / 3.0) / best / 1e9
            extra["cholesky_block_cyclic"] = {"n": n_bc, "nb": nb_bc, "ms": best * 1e3, "gflops": gf, "status": st_bc,
                                              "pct_of_fp64_peak": gf / 1e3 / (FP64_PEAK_TFLOPS * ngpus) * 100,
                                              "residual": res, "residual_bound": 10 * n_bc * eps, "residual_ok": bool(res <= 10 * n_bc * eps),
                                              "residual_def": "||A - L L^T||_F / ||A||_F over the lower triangle, evaluated on the distributed layout",
                                              "exchange": "NCCL broadcast of each factored panel from its owner; status word and pivots stay on the device"}
            best = None
            for _ in range(2):
                Abc.fill_uniform(6)
                pairs, dt = timed(lambda: lu_block_cyclic(Abc))
                best = dt if best is None else min(best, dt)
            res = lu_residual_block_cyclic(Abc, 6)
            gf = (2.0 * n_bc ** 3 / 3.0) / best / 1e9
            extra["lu_block_cyclic"] = {"n": n_bc, "nb": nb_bc, "ms": best * 1e3, "gflops": gf, "nswaps": len(pairs),
                                        "pct_of_fp64_peak": gf / 1e3 / (FP64_PEAK_TFLOPS * ngpus) * 100,
                                        "residual": res, "residual_bound": 10 * n_bc * eps, "residual_ok": bool(res <= 10 * n_bc * eps),
                                        "residual_def": "||P A - L U||_F / ||A||_F evaluated on the distributed layout",
                                        "exchange": "NCCL broadcast of each factored panel and its pivot vector from the owner"}
            del Abc
            torch.cuda.empty_cache()
            # pivots: block-cyclic over all ranks vs one GPU (na_lu_f64_dev) on the same N = 16384 matrix
            n_p = 16384
            Ap = ColumnBlockCyclic(n_p, nb_bc, rank, world, ops)
            Ap.fill_uniform(6)
            pairs_bc = lu_block_cyclic(Ap)
            same = None
            if rank == 0:
                import ctypes as C
                A1 = torch.empty(n_p * n_p, dtype=torch.float64, device=dev)
                _capi.check(L.na_fill_uniform_dev(A1.data_ptr(), n_p, n_p, n_p, 6, stream))
                sw = (C.c_size_t * (2 * n_p))(); ns = C.c_size_t(0)
                _capi.check(L.na_lu_f64_dev(n_p, n_p, A1.data_ptr(), n_p, sw, C.addressof(ns), stream))
                pairs_1 = [(sw[2 * i], sw[2 * i + 1]) for i in range(ns.value)]
                same = pairs_1 == [tuple(p) for p in pairs_bc]
                first = next((i for i, (x, y) in enumerate(zip(pairs_1, pairs_bc)) if tuple(x) != tuple(y)), None)
                extra["lu_block_cyclic"]["pivots_equal_single_gpu"] = {"n": n_p, "equal": bool(same), "npairs": len(pairs_1),
                                                                       "first_divergence": first}
                del A1
            del Ap
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
                                        "over NVLink (copy-engine peer copies from CUDA-IPC mapped buffers, or one NCCL all-gather per panel) overlapped with the K-chunked GEMM"),
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
            line["config"]["exchange"] = g2d.exchange
        if oracle_check is not None:
            line["config"]["oracle_spot_check"] = oracle_check
        if traffic_note:
            line["roofline"]["traffic_note"] = traffic_note
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    g2d.close()
    if world > 1:
        dist.destroy_process_group()


def _roof(flops, ms):
    tf = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP64_PEAK_TFLOPS,
            "note": "whole factorization against the DMMA roofline its trailing updates run on; panels are latency-bound"}


def factorization_extras(L, _capi, torch, dev, stream, N, cpu=True, e2e=True):
    """BASELINE configs[0], [2], [3], [4] on one GPU: device-resident (CUDA events), end to end through the host-pointer
    C ABI with pinned buffers, and next to the oracle's CPU time (n^3-extrapolated)."""
    import ctypes as C
    out = {}
    base = oracle_factor_baselines({"cholesky": [1024, 2048, 4096], "lu": [1024, 2048, 4096], "qr": [1024, 2048],
                                    "hessenberg": [512, 1024], "symmetric_tridiagonal": [512, 1024], "bidiagonal": [512, 1024]}) if cpu else {}

    def dev_time(fn, reps):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
        return best, r

    def host_time(fn, reps):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter(); fn(); dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        return best

    # ---- configs[0]: 1024^3 GEMM, device and host pointers
    n0 = 1024
    a = torch.empty(n0 * n0, dtype=torch.float64, device=dev); b = torch.empty_like(a); c = torch.empty_like(a)
    _capi.check(L.na_fill_uniform_dev(a.data_ptr(), n0, n0, n0, 1, stream)); _capi.check(L.na_fill_uniform_dev(b.data_ptr(), n0, n0, n0, 2, stream))
    def g0():
        for _ in range(20):
            _capi.check(L.na_dgemm_dev(n0, n0, n0, 1.0, a.data_ptr(), 1, n0, b.data_ptr(), 1, n0, 0.0, c.data_ptr(), 1, n0, stream))
    ms0, _ = dev_time(g0, 3); ms0 /= 20
    ha = a.cpu().pin_memory(); hb = b.cpu().pin_memory(); hc = torch.empty(n0 * n0, dtype=torch.float64).pin_memory()
    ms0h = host_time(lambda: _capi.check(L.na_dgemm(n0, n0, n0, 1.0, ha.data_ptr(), 1, n0, hb.data_ptr(), 1, n0, 0.0, hc.data_ptr(), 1, n0)), 5)
    out["gemm_n1024"] = {"ms": ms0, "gflops": 2.0 * n0 ** 3 / ms0 / 1e6, "pct_of_fp64_peak": 2.0 * n0 ** 3 / ms0 / 1e9 / FP64_PEAK_TFLOPS * 100,
                         "e2e": {"ms": ms0h, "gflops": 2.0 * n0 ** 3 / ms0h / 1e6, "h2d_bytes": 2 * n0 * n0 * 8, "d2h_bytes": n0 * n0 * 8,
                                 "api": "na_dgemm (host pointers, pinned)"},
                         "config": "BASELINE configs[0]: 64 output tiles on 148 SMs, launch-bound"}
    del a, b, c, ha, hb, hc

    # ---- f32 GEMM (matrixmultiply::sgemm's seam): tcgen05 kind::tf32, 3xTF32, TMEM accumulators
    a = torch.rand(N * N, dtype=torch.float32, device=dev) - 0.5; b = torch.rand(N * N, dtype=torch.float32, device=dev) - 0.5
    c = torch.empty(N * N, dtype=torch.float32, device=dev)
    ms32, _ = dev_time(lambda: _capi.check(L.na_sgemm_dev(N, N, N, 1.0, a.data_ptr(), 1, N, b.data_ptr(), 1, N, 0.0, c.data_ptr(), 1, N, stream)), 3)
    err32 = float((c.view(N, N).t()[:32].double() - a.view(N, N).t()[:32].double() @ b.view(N, N).t().double()).abs().max())
    tf32_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tf32_peak = json.load(f)["bf16_tflops"] / 2.0          # kind::tf32 runs at half the bf16 rate; no TF32 entry in the file
    except Exception:
        tf32_peak = 2250.0 / 2.0
    tf = 2.0 * N ** 3 / ms32 / 1e9
    out["sgemm_n16384"] = {"ms": ms32, "gflops": tf * 1e3, "max_abs_err_vs_f64_sample": err32,
                           "roofline": {"bound": "tensor", "achieved": 3.0 * tf, "peak": tf32_peak, "unit": "TFLOP/s", "frac": 3.0 * tf / tf32_peak,
                                        "note": "3xTF32: three kind::tf32 MMAs per f32 product, so the tensor pipe does 3x the f32-equivalent flops; "
                                                "peak = measured bf16 (MEASURED_PEAKS.json, burst) / 2; the SS-mode 128x128x8 MMA is shared-memory-bandwidth bound at ~1/2 of it"},
                           "fp32_ffma_peak_tflops": 148 * 128 * 2 * 1.965e-3,
                           "kernel": "sgemm_tcgen05_3xtf32_kernel (TMA + tcgen05.mma + TMEM), pack/split pass included"}
    del a, b, c

    # ---- configs[2]: Cholesky
    A = torch.empty(N * N, dtype=torch.float64, device=dev)
    A0 = torch.empty(N * N, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_spd_block_dev(A0.data_ptr(), N, N, N, 5, 0, 0, N, stream))       # (B+B^T)/2 + n*I, SURVEY 8(d) Cfg 3 (ii)
    fail = C.c_size_t(0)
    def chol():
        A.copy_(A0)
        return _capi.check(L.na_cholesky_f64_dev(N, A.data_ptr(), N, 0, 0.0, C.addressof(fail), stream))
    copy_ms, _ = dev_time(lambda: A.copy_(A0), 3)
    ms, st = dev_time(chol, 3); ms -= copy_ms
    fl = N ** 3 / 3.0
    out["cholesky_n16384"] = {"ms": ms, "gflops": fl / ms / 1e6, "pct_of_fp64_peak": fl / ms / 1e9 / FP64_PEAK_TFLOPS * 100, "status": st,
                              "roofline": _roof(fl, ms)}
    if e2e:
        hA = torch.empty(N * N, dtype=torch.float64).pin_memory(); hA0 = A0.cpu()
        def chol_h():
            hA.copy_(hA0)
            t0 = time.perf_counter()
            _capi.check(L.na_cholesky_f64(N, hA.data_ptr(), N, 0, 0.0, C.addressof(fail)))
            return (time.perf_counter() - t0) * 1e3
        msh = min(chol_h() for _ in range(2))
        tri_bytes = sum((N - j) * min(512, N - j) for j in range(0, N, 512)) * 8     # lower triangle as 512-column trapezoids, each way
        out["cholesky_n16384"]["e2e"] = {"ms": msh, "gflops": fl / msh / 1e6, "h2d_bytes": tri_bytes, "d2h_bytes": tri_bytes,
                                         "api": "na_cholesky_f64 (host pointer, pinned): lower triangle only, finished panels stream back behind the factorization"}
        del hA, hA0
    if "cholesky" in base:
        out["cholesky_n16384"]["cpu_baseline"] = base["cholesky"]
    del A0

    # ---- configs[3]: LU + solve with 64 right-hand sides
    A0 = torch.empty(N * N, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 6, stream))
    swaps = (C.c_size_t * (2 * N))(); ns = C.c_size_t(0)
    def lu():
        A.copy_(A0)
        _capi.check(L.na_lu_f64_dev(N, N, A.data_ptr(), N, swaps, C.addressof(ns), stream))
    ms, _ = dev_time(lu, 3); ms -= copy_ms
    fl = 2.0 * N ** 3 / 3.0
    out["lu_n16384"] = {"ms": ms, "gflops": fl / ms / 1e6, "pct_of_fp64_peak": fl / ms / 1e9 / FP64_PEAK_TFLOPS * 100, "nswaps": ns.value,
                        "roofline": _roof(fl, ms)}
    nrhs = 64
    B0 = torch.empty(N * nrhs, dtype=torch.float64, device=dev); X = torch.empty_like(B0)
    _capi.check(L.na_fill_uniform_dev(B0.data_ptr(), N, nrhs, N, 7, stream))
    def solve():
        X.copy_(B0)
        return _capi.check(L.na_lu_solve_f64_dev(N, A.data_ptr(), N, swaps, ns.value, X.data_ptr(), N, nrhs, stream))
    ms_s, st_s = dev_time(solve, 3)
    # residual of the solve: ||A x - b||_inf / (||A||_inf ||x||_inf) with our own GEMM
    R = B0.clone()
    _capi.check(L.na_dgemm_dev(N, N, nrhs, 1.0, A0.data_ptr(), 1, N, X.data_ptr(), 1, N, -1.0, R.data_ptr(), 1, N, stream))
    torch.cuda.synchronize()
    rs = float(R.abs().max() / (A0.view(N, N).abs().sum(dim=0).max() * X.abs().max()))
    out["lu_n16384"]["solve_64rhs"] = {"ms": ms_s, "gflops": 2.0 * N * N * nrhs / ms_s / 1e6, "status": st_s, "scaled_residual": rs,
                                       "api": "na_lu_solve_f64_dev (LU::solve_mut, lu.rs:242-260)"}
    del B0, X, R
    if e2e:
        hA = torch.empty(N * N, dtype=torch.float64).pin_memory(); hA0 = A0.cpu()
        def lu_h():
            hA.copy_(hA0)
            t0 = time.perf_counter()
            _capi.check(L.na_lu_f64(N, N, hA.data_ptr(), N, swaps, C.addressof(ns)))
            return (time.perf_counter() - t0) * 1e3
        msh = min(lu_h() for _ in range(2))
        out["lu_n16384"]["e2e"] = {"ms": msh, "gflops": fl / msh / 1e6, "h2d_bytes": N * N * 8, "d2h_bytes": N * N * 8,
                                   "api": "na_lu_f64 (host pointer, pinned)"}
        del hA, hA0
    if "lu" in base:
        out["lu_n16384"]["cpu_baseline"] = base["lu"]
    del A, A0

    # ---- configs[4]: Householder QR, tall-skinny
    m, n = 65536, 4096
    A = torch.empty(m * n, dtype=torch.float64, device=dev)
    A0 = torch.empty(m * n, dtype=torch.float64, device=dev)
    diag = torch.empty(n, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), m, n, m, 8, stream))
    def qr():
        A.copy_(A0)
        _capi.check(L.na_qr_f64_dev(m, n, A.data_ptr(), m, diag.data_ptr(), stream))
    ms, _ = dev_time(qr, 4); ms -= copy_ms      # the first repetition grows the stream-ordered pool
    fl = 2.0 * m * n * n - 2.0 * n ** 3 / 3.0
    out["qr_65536x4096"] = {"ms": ms, "gflops": fl / ms / 1e6, "pct_of_fp64_peak": fl / ms / 1e9 / FP64_PEAK_TFLOPS * 100, "roofline": _roof(fl, ms)}
    if e2e:
        hA = torch.empty(m * n, dtype=torch.float64).pin_memory(); hA0 = A0.cpu(); hd = torch.empty(n, dtype=torch.float64)
        def qr_h():
            hA.copy_(hA0)
            t0 = time.perf_counter()
            _capi.check(L.na_qr_f64(m, n, hA.data_ptr(), m, hd.data_ptr()))
            return (time.perf_counter() - t0) * 1e3
        msh = min(qr_h() for _ in range(2))
        out["qr_65536x4096"]["e2e"] = {"ms": msh, "gflops": fl / msh / 1e6, "h2d_bytes": m * n * 8, "d2h_bytes": m * n * 8 + n * 8,
                                       "api": "na_qr_f64 (host pointer, pinned): finished outer panels are converted and streamed back behind the factorization"}
        del hA, hA0
    if "qr" in base:
        out["qr_65536x4096"]["cpu_baseline"] = base["qr"]
    del A, A0

    # ---- FullPivLU / ColPivQR -- one pass over the trailing matrix per step (Level-2, memory-bound)
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = json.load(f)["hbm_gbs"]; peak_src = "measured (MEASURED_PEAKS.json, copy bandwidth)"
    except Exception:
        hbm_peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    n4 = 8192                                   # 537 MB: the trailing matrix does not sit in the 126 MB L2 for most of the run
    A0 = torch.empty(n4 * n4, dtype=torch.float64, device=dev); A = torch.empty_like(A0); dg = torch.empty(n4, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), n4, n4, n4, 6, stream))
    ps = (C.c_size_t * (2 * n4))(); qs = (C.c_size_t * (2 * n4))(); n_p = C.c_size_t(0); n_q = C.c_size_t(0)
    def fplu():
        A.copy_(A0)
        _capi.check(L.na_full_piv_lu_f64_dev(n4, n4, A.data_ptr(), n4, ps, C.addressof(n_p), qs, C.addressof(n_q), stream))
    def cpqr():
        A.copy_(A0)
        _capi.check(L.na_col_piv_qr_f64_dev(n4, n4, A.data_ptr(), n4, dg.data_ptr(), ps, C.addressof(n_p), stream))
    alg_bytes = 16.0 * n4 ** 3 / 3.0            # every step reads and writes the (n - i)^2 trailing matrix once
    for key, fn, what in (("full_piv_lu_n8192", fplu, "full_piv_lu_kernel"), ("col_piv_qr_n8192", cpqr, "col_piv_qr_kernel")):
        ms, _ = dev_time(fn, 2)
        gbs = alg_bytes / ms / 1e6
        out[key] = {"ms": ms, "us_per_step": ms * 1e3 / n4,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                                 "kernel": what, "peak_source": peak_src,
                                 "algorithmic_bytes_per_launch": alg_bytes,
                                 "note": "16 n^3 / 3 bytes: one read + one write of the trailing matrix per elimination step "
                                         "(ColPivQR reads every column twice: dot product, then reflection); "
                                         "two grid barriers per step bound the small-n end"}}
    del A, A0

    # ---- Hessenberg / SymmetricTridiagonal / Bidiagonal -- one fused product pass + one update pass over the
    # trailing block per step (Bidiagonal: two product passes); n = 8192 so that the matrix (537 MB) does not sit in the L2
    n8 = 8192
    A0 = torch.empty(n8 * n8, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
    dg = torch.empty(n8, dtype=torch.float64, device=dev); eg = torch.empty(n8, dtype=torch.float64, device=dev)
    _capi.check(L.na_fill_uniform_dev(A0.data_ptr(), n8, n8, n8, 6, stream))
    def hess():
        A.copy_(A0)
        _capi.check(L.na_hessenberg_f64_dev(n8, A.data_ptr(), n8, dg.data_ptr(), stream))
    def symtri():
        A.copy_(A0)
        _capi.check(L.na_symmetric_tridiagonal_f64_dev(n8, A.data_ptr(), n8, dg.data_ptr(), stream))
    def bidiag():
        A.copy_(A0)
        _capi.check(L.na_bidiagonal_f64_dev(n8, n8, A.data_ptr(), n8, dg.data_ptr(), eg.data_ptr(), stream))
    for key, fn, what, alg_bytes, flops, note in (
            ("hessenberg_n8192", hess, "two_sided_fused_kernel<false>", 8.0 * n8 ** 3, 10.0 * n8 ** 3 / 3.0,
             "8 n^3 bytes: per step ONE read + ONE write of the n x (n - k) block (the update of step s - 1 and the products of "
             "step s, w = A u and z = A^T u, in the same pass; n >= 6144).  Against the 12 n^3 of the two-pass kernel the same "
             "time reads as 1.5x this bandwidth"),
            ("symmetric_tridiagonal_n8192", symtri, "two_sided_kernel<true>", 4.0 * n8 ** 3, 4.0 * n8 ** 3 / 3.0,
             "4 n^3 bytes: per step one read + one read-modify-write of the lower triangle of the (n - k)^2 block"),
            ("bidiagonal_n8192", bidiag, "bidiagonal_kernel", 32.0 * n8 ** 3 / 3.0, 8.0 * n8 ** 3 / 3.0,
             "32 n^3 / 3 bytes: per step two reads (column products, row products of the column-reflected block) + one read-modify-write")):
        ms, _ = dev_time(fn, 2)
        gbs = alg_bytes / ms / 1e6
        out[key] = {"ms": ms, "us_per_step": ms * 1e3 / n8, "gflops": flops / ms / 1e6,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                                 "kernel": what, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "note": note}}
        if key.rsplit("_", 1)[0] in base:
            out[key]["cpu_baseline"] = base[key.rsplit("_", 1)[0]]
    if e2e:
        # the same three through the host-pointer entry points (pinned host matrix; upload, reduce, download: the first step
        # touches every element and every step rewrites the trailing block, so nothing can stream)
        try:
            hA = torch.empty(n8 * n8, dtype=torch.float64).pin_memory(); hA0 = A0.cpu()
            hd = torch.empty(n8, dtype=torch.float64); he = torch.empty(n8, dtype=torch.float64)
            for key, call, api in (
                    ("hessenberg_n8192", lambda: L.na_hessenberg_f64(n8, hA.data_ptr(), n8, hd.data_ptr()), "na_hessenberg_f64"),
                    ("symmetric_tridiagonal_n8192", lambda: L.na_symmetric_tridiagonal_f64(n8, hA.data_ptr(), n8, hd.data_ptr()), "na_symmetric_tridiagonal_f64"),
                    ("bidiagonal_n8192", lambda: L.na_bidiagonal_f64(n8, n8, hA.data_ptr(), n8, hd.data_ptr(), he.data_ptr()), "na_bidiagonal_f64")):
                hA.copy_(hA0)
                t0 = time.perf_counter()
                _capi.check(call())
                msh = (time.perf_counter() - t0) * 1e3
                out[key]["e2e"] = {"ms": msh, "h2d_bytes": n8 * n8 * 8, "d2h_bytes": n8 * n8 * 8 + 2 * n8 * 8,
                                   "api": api + " (host pointer, pinned)"}
            del hA, hA0
        except Exception as ex:
            out["hessenberg_n8192"]["e2e"] = {"error": repr(ex)}
    try:   # DRAM bytes of the Hessenberg launch from the committed ncu capture, valid for the kernel source it was taken with
        import hashlib
        with open(os.path.join(ROOT, "profiles", "twosided_traffic.json")) as f:
            tj = json.load(f)
        with open(os.path.join(ROOT, "nalgebra_b200", "csrc", "factor_twosided.cu"), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() == tj.get("kernel_source_sha256"):
                out["hessenberg_n8192"]["roofline"]["traffic"] = tj["traffic_bytes_per_launch"]
    except Exception:
        pass
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
