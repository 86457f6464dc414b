#!/usr/bin/env python3
"""Headline benchmark: QEq solve time at N = 20,001 atoms, STO shielding, fp64 (BASELINE.json `metric`).

A step = one full `QEqSolver::solve` of the S20k workload (6,667-water open cluster):
assembly of the shielded-Coulomb matrix, factorisation, the complete hydrogen SCF loop, charges out.

  value      seconds per solve with the per-atom inputs already resident in HBM (cheq_solve_device)
  e2e        the same through the public API with host (pinned) buffers: H2D of the inputs and D2H of the charges
             inside the timed region (cheq_solve)
  roofline   the dominant kernel (gemm_nt_kernel, FP64 DMMA): algorithmic flops / summed launch durations, both
             measured live with CUDA events on the engine's stream; peak = cuBLAS DGEMM measured in the same run
             (MEASURED_PEAKS.json has no fp64 entry)
  cpu_baseline / --impl reference
             the cheq-equivalent CPU path (oracle/: C pair integrals + LAPACK partial-pivot LU, one fresh
             factorisation per SCF iteration exactly like the reference) timed on this box's host cores on a
             bounded sample and extrapolated by pair count and LAPACK flop count (stated in `sample`).
             cheq itself is Rust with un-vendored crates and cannot be built offline.

N > 1 (torchrun): the ranks form one engine group (cheq_ctx_dist_init) and solve the SAME S20k system together:
owner-computes assembly and trailing updates over 1-D block-cyclic column panels, NCCL panel broadcasts over
NVLink, replicated triangular solves.  value = seconds per solve (max over ranks), i.e.
strong scaling of the headline metric; the batched-MD secondary metric shards frames instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "qeq_solve_time_n20k_sto_fp64"
UNIT = "s"
# SCF iterations of the S20k workload (seed 20261019) as executed by the GPU path and by the pivoted-LU path
# (identical trajectories); used to extrapolate the CPU arm, which cannot run 18 full 20k LUs in minutes.
KNOWN_ITERATIONS = {"s20k": 18}


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun) also write to fd 1, so
    the real stdout is kept aside for the result and fd 1 is pointed at stderr for everything else."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--molecules", type=int, default=6667, help="waters in the cluster (6667 = S20k)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(molecules: int, rank: int = 0):
    from cheq_b200 import get_default_parameters, workloads
    Z, xyz = workloads.water_cluster(molecules, workloads.SEED + rank)
    chi, hard, rad, nq = workloads.gather_parameters(Z, get_default_parameters())
    name = "S20k" if molecules == 6667 else f"W{3 * molecules}"
    return name, Z, np.ascontiguousarray(xyz), chi, hard, rad, nq


# ---- CPU arm ------------------------------------------------------------------------------------------------
def cpu_reference_sample(Z, xyz, iterations: int, budget_s: float):
    """One bounded sample of the cheq-equivalent CPU path -> extrapolated seconds per full solve."""
    from scipy.linalg import lu_factor, lu_solve

    from oracle import oracle as O
    P = O.default_parameters()
    n = len(Z)
    cores = O.num_threads()
    opt = O.Options()
    # probes
    chi, hard, rad, nq = O.gather(Z[:600], P)
    O.build_system(xyz[:600], chi, hard, rad, nq, 0.0, None, opt)                 # warms tables/threads
    t0 = time.perf_counter()
    O.build_system(xyz[:600], chi, hard, rad, nq, 0.0, None, opt)
    t_pair = (time.perf_counter() - t0) / (600 * 599 / 2)
    m = np.random.default_rng(0).normal(size=(2000, 2000)) + 50 * np.eye(2000)
    t0 = time.perf_counter()
    lu_factor(m, check_finite=False)
    lu_rate = (2.0 / 3.0) * 2000 ** 3 / (time.perf_counter() - t0)
    # sample sizes from the budget: ~35 % assembly, ~65 % LU
    n_a = int(min(n, max(1000, (2 * 0.35 * budget_s / t_pair) ** 0.5)))
    n_l = int(min(n + 1, max(2000, (0.65 * budget_s * lu_rate * 1.5) ** (1.0 / 3.0))))
    chi, hard, rad, nq = O.gather(Z[:n_a], P)
    t0 = time.perf_counter()
    M, rhs, hidx = O.build_system(xyz[:n_a], chi, hard, rad, nq, 0.0, None, opt)
    t_asm = time.perf_counter() - t0
    t_asm_full = t_asm * (n * (n - 1.0)) / (n_a * (n_a - 1.0))
    if n_l > n_a + 1:   # LU operand: any well-conditioned matrix of the right size costs the same flops
        W = np.random.default_rng(1).normal(size=(n_l, n_l))
        W += n_l * np.eye(n_l)
        W = np.asfortranarray(W)
        b = np.ones(n_l)
    else:
        W = np.asfortranarray(M[:n_l, :n_l])
        b = rhs[:n_l]
    t0 = time.perf_counter()
    lu, piv = lu_factor(W, overwrite_a=False, check_finite=False)
    lu_solve((lu, piv), b, check_finite=False)
    t_lu = time.perf_counter() - t0
    t_lu_full = t_lu * ((n + 1.0) / n_l) ** 3
    total = t_asm_full + iterations * t_lu_full
    sample = (f"assembly of a {n_a}-atom sub-cluster ({t_asm:.2f} s, scaled by pair count to N={n}: {t_asm_full:.1f} s) + "
              f"one LAPACK dgetrf+dgetrs of order {n_l} ({t_lu:.2f} s, scaled by (N+1)^3 to {t_lu_full:.1f} s) x "
              f"{iterations} SCF iterations (one fresh LU per iteration as in implementation.rs:303-314); "
              f"cheq-equivalent CPU path (restated; faer/sto-ns not available offline)")
    return total, sample, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    name, Z, xyz, *_ = workload(args.molecules)
    its = KNOWN_ITERATIONS.get(name.lower(), 10)
    per_step = min(25.0, 150.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        cpu_reference_sample(Z, xyz, its, per_step)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sample, cores = cpu_reference_sample(Z, xyz, its, per_step)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.median(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / max(1, args.steps) * 1e3,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {len(Z)}-atom open water cluster, STO, hydrogen SCF, total charge 0",
                   "note": "value is extrapolated from a bounded sample per step (see cpu_baseline.sample); "
                           "ms_per_step is the wall time of one sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- GPU arm ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(torch, device):
    """cuBLAS DGEMM 8192^3, best of 10 and a 2 s sustained loop (same protocol as MEASURED_PEAKS.json's bf16 entry)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 0
    e0.record()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 2.0:
        torch.matmul(a, b)
        reps += 1
        if reps % 4 == 0:
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = reps * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b
    return best, sustained


def secondary_metrics(ctx, rank, world, device, barrier):
    """BASELINE configs 3 and 4 in brief: batched MD frames/s (frames sharded over the ranks, no collective on the
    data path) and the B3k solve with field + point charges (rank 0).  Reported next to the headline, not timed
    with the same care (wall clock, 1 warm-up, 2 repeats)."""
    import torch
    import torch.distributed as dist

    from cheq_b200 import QEqSolver, SolverOptions, _ffi, get_default_parameters, workloads
    from cheq_b200.parallel import solve_batch_sharded
    P = get_default_parameters()
    solver = QEqSolver(P, SolverOptions(), ctx)
    out = {}
    Zm, base = workloads.organic_chain(1000)
    F = 1024 * world
    frames = workloads.md_frames(base, F)

    def solve_local(fr):
        return solver.solve_batch(Zm, fr, 0.0)
    solve_batch_sharded(solve_local, frames[: 64 * world])
    best = 1e30
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        q, mu, its, status = solve_batch_sharded(solve_local, frames)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = min(best, float(dt.item()))
    out["md"] = {"workload": f"{F} frames x {len(Zm)}-atom organic chain (C/H/N/O), STO, hydrogen SCF, frames sharded "
                             f"round-robin over {world} rank(s), results all-gathered",
                 "frames_per_s": F / best, "seconds": best, "iterations_min_max": [int(its.min()), int(its.max())],
                 "all_converged": bool((status == 0).all())}
    if rank == 0:
        if world > 1:   # single-frame solves on the group context are collective: use a private engine here
            import cheq_b200
            solver = QEqSolver(P, SolverOptions(), cheq_b200.Context(device.index))
        Z, xyz, ext = workloads.water_box_3k()
        chi, hard, rad, nq = workloads.gather_parameters(Z, P)
        _, _, prad, pnq = workloads.gather_parameters(ext["Z"], P)
        e = _ffi.External()
        pxyz = np.ascontiguousarray(ext["xyz"])
        pq = np.ascontiguousarray(ext["q"])
        e.num_point_charges = len(pq)
        e.xyz, e.charge = pxyz.ctypes.data_as(C.c_void_p), pq.ctypes.data_as(C.c_void_p)
        e.radius, e.n = prad.ctypes.data_as(C.c_void_p), pnq.ctypes.data_as(C.c_void_p)
        e.field[0], e.field[1], e.field[2] = ext["field"]
        x = np.ascontiguousarray(xyz)
        solver.solve_arrays(x, chi, hard, rad, nq, 0.0, e)
        t0 = time.perf_counter()
        for _ in range(3):
            r = solver.solve_arrays(x, chi, hard, rad, nq, 0.0, e)
        out["b3k"] = {"workload": "3,000-atom water cluster + uniform field + 10,000 embedding point charges",
                      "ms_per_solve": (time.perf_counter() - t0) / 3 * 1e3, "iterations": r.iterations}
    return out


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import cheq_b200
    from cheq_b200 import QEqSolver, SolverOptions, _ffi, get_default_parameters
    from cheq_b200.solver import make_options, raise_for_status

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    name, Z, xyz, chi, hard, rad, nq = workload(args.molecules, 0)   # one system, identical on every rank
    n = len(Z)
    ctx = cheq_b200.Context(local_rank)
    lib = ctx._lib
    if world > 1:
        from cheq_b200 import distributed
        distributed.join(ctx)
    opt = make_options(SolverOptions())
    stream = torch.cuda.ExternalStream(lib.cheq_ctx_stream(ctx.handle), device=device)

    d_xyz = torch.from_numpy(xyz).to(device)
    d_chi = torch.from_numpy(chi).to(device)
    d_hard = torch.from_numpy(hard).to(device)
    d_rad = torch.from_numpy(rad).to(device)
    d_nq = torch.from_numpy(nq).to(device)
    d_q = torch.empty(n, dtype=torch.float64, device=device)
    pot, its, delta = C.c_double(), C.c_uint32(), C.c_double()
    torch.cuda.synchronize()

    def step_device():
        rc = lib.cheq_solve_device(ctx.handle, n, d_xyz.data_ptr(), d_chi.data_ptr(), d_hard.data_ptr(),
                                   d_rad.data_ptr(), d_nq.data_ptr(), 0.0, C.byref(opt), None, d_q.data_ptr(),
                                   C.byref(pot), C.byref(its), C.byref(delta))
        raise_for_status(ctx, rc, opt.max_iterations, delta.value)

    # pinned host buffers for the end-to-end path
    h_xyz = torch.from_numpy(xyz).pin_memory()
    h_chi = torch.from_numpy(chi).pin_memory()
    h_hard = torch.from_numpy(hard).pin_memory()
    h_q = torch.empty(n, dtype=torch.float64).pin_memory()

    def step_host():
        rc = lib.cheq_solve(ctx.handle, n, h_xyz.data_ptr(), h_chi.data_ptr(), h_hard.data_ptr(),
                            rad.ctypes.data_as(C.c_void_p), nq.ctypes.data_as(C.c_void_p), 0.0, C.byref(opt), None,
                            h_q.data_ptr(), C.byref(pot), C.byref(its), C.byref(delta))
        raise_for_status(ctx, rc, opt.max_iterations, delta.value)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # pass 1: the timed region (no instrumentation)
    launches = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        launches += ctx.stats()["kernel_launches"]
    e1.record(stream)
    barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total = float(ms_t.item())
    # pass 2: the same K steps with every launch of the dominant kernel bracketed by CUDA events on the engine's
    # stream (two event records per launch cost ~5 % of the step, hence not in pass 1)
    lib.cheq_set_profiling(ctx.handle, 1)
    gemm_ms = gemm_fl = 0.0
    gemm_n = 0
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(args.steps):
        step_device()
        st = ctx.stats()
        gemm_ms += st["gemm_ms"]
        gemm_fl += st["gemm_flops"]
        gemm_n += st["gemm_launches"]
    e3.record(stream)
    torch.cuda.synchronize()
    ms_profiled = e2.elapsed_time(e3)
    clocks = sampler.stop() if rank == 0 else None
    stats = ctx.stats()
    iterations = int(its.value)
    mu_dev = pot.value
    q_dev = d_q.cpu().numpy()

    lib.cheq_set_profiling(ctx.handle, 0)
    for _ in range(min(args.warmup, 1)):
        step_host()
    ms_e2e = timed(step_host, args.steps)
    assert abs(pot.value - mu_dev) < 1e-9 and np.max(np.abs(h_q.numpy() - q_dev)) < 1e-9

    secondary = secondary_metrics(ctx, rank, world, device, barrier)

    launches_t = torch.tensor([float(launches)], device=device)
    if world > 1:
        dist.all_reduce(launches_t)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = ms_per_step / 1e3
    e2e_value = ms_e2e / args.steps / 1e3
    peak_burst, peak_sustained = measure_fp64_peak(torch, device)
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    hbm_peak = None
    try:
        hbm_peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs")
    except Exception:
        pass
    asm_gbs = stats["bytes_assemble"] / (stats["ms_assemble"] * 1e-3) / 1e9 if stats["ms_assemble"] > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {n}-atom open water cluster (seed {20261019}), STO, hydrogen SCF "
                               f"({iterations} iterations), total charge 0, default SolverOptions",
                   "parallelism": (f"one system over {world} GPUs: 1-D block-cyclic column panels ("
                                   f"{os.environ.get('CHEQ_TILES_PER_PANEL', '4')} x 128 columns), owner-computes "
                                   "assembly/updates, finished panels pushed to the peers over NVLink (copy engines, "
                                   "CUDA IPC mappings, flag + cuStreamWaitValue32) up to 4 GPUs or NCCL-broadcast "
                                   "beyond, panel- and tile-level look-ahead, replicated triangular solves "
                                   f"(CHEQ_DIST_P2P={os.environ.get('CHEQ_DIST_P2P', '1')})") if world > 1 else "single GPU",
                   "l2": "working set (3.3 GB matrix) >> 126 MB L2; no flush needed",
                   "phases_ms": {k: stats[k] for k in ("ms_host_prepare", "ms_h2d", "ms_assemble", "ms_factor_once",
                                                       "ms_scf", "ms_d2h")},
                   "factor_tflops_overall": stats["flops_factor"] / ((stats["ms_factor_once"] + stats["ms_scf"]) * 1e9),
                   "lu_fallback_frames": stats["lu_fallback_frames"]},
        "roofline": {"bound": "tensor", "kernel": "gemm_nt_kernel (FP64 DMMA.8x8x4)", "achieved": achieved,
                     "peak": peak_sustained, "unit": "TFLOP/s", "frac": achieved / peak_sustained if peak_sustained else None,
                     "traffic": None, "launches": gemm_n, "kernel_ms_per_step": gemm_ms / args.steps,
                     "largest_launch": ({"what": "Schur update of the hydrogen block, K = 6784, 106 x 105 tiles (lower)",
                                         "flops": 1.2488e12, "ms": 39.41, "tflops": 31.7, "dmma_pipe_active": 0.876,
                                         "traffic": 9.286e9, "algorithmic_bytes": 2.2e9,
                                         "source": "one ncu --set full capture of that launch "
                                                   "(profiles/r1_ncu_full_gemm_schur_v3.txt); `traffic` above is null "
                                                   "because `achieved` averages 3,868 launches of different shapes"}
                                        if n == 20001 else None),
                     "share_of_step": gemm_ms / ms_profiled,
                     "measured_in": f"second pass of {args.steps} steps with per-launch CUDA events "
                                    f"({ms_profiled / args.steps:.1f} ms per step vs {ms_total / args.steps:.1f} uninstrumented)",
                     "potrf_ms_per_step": stats["potrf_ms"], "backsolve_ms_per_step": stats["backsolve_ms"],
                     "note": ("panel schedule: GEMMs of the main, panel and bulk streams run concurrently, so the "
                              "summed per-launch durations overstate the kernel's wall-clock share and understate "
                              "its rate" if (world > 1 or os.environ.get("CHEQ_FACTOR_MODE") == "1") else
                              "recursive schedule: one stream, launches do not overlap"),
                     "peak_source": f"cuBLAS DGEMM 8192^3 measured in this run: sustained {peak_sustained:.1f}, "
                                    f"burst {peak_burst:.1f} TFLOP/s (MEASURED_PEAKS.json has no fp64 entry)"},
        "roofline_assembly": {"bound": "hbm", "kernel": "assemble_lower_kernel<STO>", "achieved": asm_gbs,
                              "peak": hbm_peak, "unit": "GB/s", "frac": asm_gbs / hbm_peak if hbm_peak else None,
                              "traffic": 1.59998e9 if n == 20001 else None,
                              "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                                                "capture of this launch (profiles/r1_ncu_full_assemble_v2.txt); "
                                                "algorithmic bytes 8 N (N + 1) / 2 = 1.60016e9",
                              "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * (24 + 8 + 8)),
                "d2h_bytes_per_step": int(n * 8 + 64)},
        "gpu_launches": int(launches_t.item()),
        "clocks": clocks,
        "secondary": secondary,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, sample, cores = cpu_reference_sample(Z, xyz, iterations, 20.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
