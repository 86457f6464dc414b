#!/usr/bin/env python3
"""Headline benchmark: QEq solve time at N = 20,001 atoms, STO shielding, fp64 (BASELINE.json `metric`).

A step = one full `QEqSolver::solve` of the S20k workload (6,667-water open cluster):
assembly of the shielded-Coulomb matrix, factorisation, the complete hydrogen SCF loop, charges out.

  value      seconds per solve with the per-atom inputs already resident in HBM (cheq_solve_device)
  e2e        the same through the public API with host (pinned) buffers: H2D of the inputs and D2H of the charges
             inside the timed region (cheq_solve)
  roofline   the dominant kernel, measured live with CUDA events around every launch on the engine's stream (second pass):
             oz_gemm_kernel (the deep FP64 updates as int8 slice products on tcgen05): int8 tensor operations executed /
             summed launch durations incl. the slicing passes, against the int8 tensor peak (2 x MEASURED_PEAKS.json's
             bf16 figure, or the higher int8 issue rate measured by tools/tc05_probe); its FP64-equivalent rate is
             reported beside it.  `roofline_other_gemm` is gemm_nt_kernel (FP64 DMMA: the shallow updates) against cuBLAS
             DGEMM measured in the same run (MEASURED_PEAKS.json has no fp64 entry); with the tcgen05 path off it is the
             dominant kernel and takes the `roofline` slot.  `roofline_matvec` is the HBM-bound kernel of the stale-factor
             SCF iterations (tile_matvec_kernel), `roofline_assembly` the assembly kernel.
  parity     the timed result against ONE real full solve of the same workload by the CPU oracle (the reference's
             algorithm as written; cached under /tmp, computed by the reference arm or, if absent, here)
  cpu_baseline / --impl reference
             the cheq-equivalent CPU path (oracle/: C pair integrals + LAPACK partial-pivot LU, one fresh
             factorisation per SCF iteration exactly like the reference) on all host cores of this box: one REAL full
             solve (measured, cached) + K timed samples, each a full-size LU + solve of the real bordered matrix;
             value = assembly + iterations x median(sample).  cheq itself is Rust with un-vendored crates and cannot
             be built offline (kind "port").

N > 1 (torchrun): the ranks form one engine group (cheq_ctx_dist_init) and solve the SAME S20k system together
(strong scaling): owner-computes assembly and trailing updates over 1-D block-cyclic column panels, panel exchange
over NVLink, replicated triangular solves.  Before timing, rank 0 solves the system once on a
private single-GPU engine and the collective result must match it to 1e-10.

--workload md     BASELINE config 4: 10,000 frames x 1,000-atom organic molecule, frames sharded over the ranks
                  (weak scaling in frames per GPU is NOT used: the trajectory is fixed, value = frames / s)
--workload s100k  BASELINE config 5: 100,002-atom water cluster, one system over the ranks
--workload b3k    BASELINE config 3: 3,000 atoms + uniform field + 10,000 point charges (one GPU)
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

UNIT = "s"
SEED = 20261019
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
    ap.add_argument("--workload", default="s20k", choices=["s20k", "md", "s100k", "b3k"])
    ap.add_argument("--molecules", type=int, default=0, help="waters in the cluster (default: 6667 = S20k, 33334 = S100k)")
    ap.add_argument("--frames", type=int, default=10000, help="md: frames of the trajectory")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (cpu_baseline, parity) unless cached")
    ap.add_argument("--no-secondary", action="store_true")
    return ap.parse_args()


def metric_name(args):
    return {"s20k": "qeq_solve_time_n20k_sto_fp64", "s100k": "qeq_solve_time_n100k_sto_fp64",
            "b3k": "qeq_solve_time_b3k_field_sto_fp64", "md": "qeq_md_frames_per_s_1000atoms_sto_fp64"}[args.workload]


def n_molecules(args):
    if args.molecules:
        return args.molecules
    return 33334 if args.workload == "s100k" else 6667


def workload(molecules: int):
    from cheq_b200 import get_default_parameters, workloads
    Z, xyz = workloads.water_cluster(molecules, workloads.SEED)
    chi, hard, rad, nq = workloads.gather_parameters(Z, get_default_parameters())
    name = {6667: "S20k", 33334: "S100k"}.get(molecules, f"W{3 * molecules}")
    return name, Z, np.ascontiguousarray(xyz), chi, hard, rad, nq


def workload_text(name, n, iterations=None):
    its = f" ({iterations} iterations)" if iterations else ""
    return (f"{name}: {n}-atom open water cluster (seed {SEED}), STO, hydrogen SCF{its}, total charge 0, "
            "default SolverOptions")


# ---- CPU arm: the oracle on all host cores ------------------------------------------------------------------
def _cache_path(name, n):
    return Path(os.environ.get("CHEQ_ORACLE_CACHE", "/tmp")) / f"cheq_b200_oracle_{name}_{SEED}_{n}.npz"


def load_oracle_result(name, n):
    p = _cache_path(name, n)
    if not p.exists():
        return None
    try:
        d = np.load(p, allow_pickle=False)
        return {k: d[k] for k in d.files}
    except Exception:
        return None


def oracle_full_solve(name, Z, xyz):
    """ONE real full solve of the workload by the CPU oracle, timed; cached under /tmp keyed by workload and seed."""
    cached = load_oracle_result(name, len(Z))
    if cached is not None:
        return cached, False
    from oracle import oracle as O
    cores = O.use_all_cores()
    opt = O.Options()
    P = O.default_parameters()
    t0 = time.perf_counter()
    chi, hard, rad, nq = O.gather(Z, P)
    base, rhs, hidx = O.build_system(xyz, chi, hard, rad, nq, 0.0, None, opt)
    t_asm = time.perf_counter() - t0
    lu_times = []

    def timed_lu(work, b):
        t1 = time.perf_counter()
        x = O._lu_solve(work, b)
        lu_times.append(time.perf_counter() - t1)
        return x
    res = O.run_scf(len(Z), base, rhs, hidx, hard, opt, lu_solve=timed_lu)
    t_total = time.perf_counter() - t0
    out = {"charges": res.charges, "mu": np.float64(res.equilibrated_potential),
           "iterations": np.int64(res.iterations), "deltas": np.array(res.deltas), "t_total": np.float64(t_total),
           "t_asm": np.float64(t_asm), "t_lu": np.array(lu_times), "cores": np.int64(cores)}
    try:
        p = _cache_path(name, len(Z))
        tmp = p.with_suffix(".tmp.npz")
        np.savez(tmp, **out)
        os.replace(tmp, p)
    except Exception:
        pass
    return out, True


def run_reference(args, rank, world):
    """--impl reference: the CPU path on this box's cores.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as O
    if args.workload == "md":
        return run_reference_md(args)
    if args.workload == "b3k":
        return run_reference_b3k(args)
    name, Z, xyz, *_ = workload(n_molecules(args))
    n = len(Z)
    cores = O.use_all_cores()
    if args.workload == "s100k":
        emit({"impl": "reference", "metric": metric_name(args), "unavailable":
              "the CPU path needs 3 x 80 GB matrices and ~10^15 flops per LU at 100k atoms: not run"})
        return
    full, fresh = oracle_full_solve(name, Z, xyz)
    its = int(full["iterations"])
    # K timed samples: each one full-size partial-pivot LU + solve of the REAL bordered matrix
    opt = O.Options()
    chi, hard, rad, nq = O.gather(Z, O.default_parameters())
    base, rhs, hidx = O.build_system(xyz, chi, hard, rad, nq, 0.0, None, opt)
    t_lu_est = float(np.median(full["t_lu"]))
    budget = 240.0
    order = n + 1
    total_samples = max(1, args.steps + args.warmup)
    if t_lu_est * total_samples > budget:      # shrink the sample, never the number of steps
        order = int(max(2000, (n + 1) * (budget / (t_lu_est * total_samples)) ** (1.0 / 3.0)))
    W = base if order == n + 1 else np.asfortranarray(base[:order, :order])
    b = rhs[:order]
    for _ in range(args.warmup):
        O._lu_solve(W, b)
    samples = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        O._lu_solve(W, b)
        samples.append((time.perf_counter() - t1) * ((n + 1.0) / order) ** 3)
    wall = time.perf_counter() - t0
    value = float(full["t_asm"]) + its * float(np.median(samples))
    sample = (f"{args.steps} samples, each one LAPACK dgetrf+dgetrs of order {order}"
              + ("" if order == n + 1 else f" (scaled by ((N+1)/{order})^3)") +
              f" on the real bordered matrix (median {np.median(samples):.2f} s) x {its} SCF iterations (one fresh LU "
              f"per iteration, implementation.rs:303-314) + assembly {float(full['t_asm']):.1f} s; ONE complete solve "
              f"was also run and timed {'in this run' if fresh else 'earlier on this box (cached)'}: "
              f"{float(full['t_total']):.1f} s wall, {its} iterations; cheq-equivalent CPU path (restated; faer/sto-ns "
              "not available offline)")
    emit({
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / max(1, args.steps) * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(name, n, its), "measured_full_solve_s": float(full["t_total"]),
                   "note": "value = assembly + iterations x median(sampled full-size LU); ms_per_step is the wall time "
                           "of one sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def run_reference_md(args):
    from oracle import oracle as O
    from cheq_b200 import workloads
    cores = O.use_all_cores()
    Zm, base = workloads.organic_chain(1000)
    per = max(1, min(4, int(120.0 / max(1, args.steps + args.warmup) / 1.5)))
    frames = workloads.md_frames(base, per)
    for _ in range(min(args.warmup, 1)):
        O.solve(Zm, frames[0])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for f in range(per):
            O.solve(Zm, frames[f])
    wall = time.perf_counter() - t0
    value = per * args.steps / wall
    emit({"impl": "reference", "metric": metric_name(args), "value": value, "unit": "frames/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": f"MD: {args.frames} frames x 1000-atom organic chain; each step solves {per} frames"},
          "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                           "sample": f"{per} frames per step, frames solved one after the other on all cores"},
          "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


def run_reference_b3k(args):
    from oracle import oracle as O
    from cheq_b200 import workloads
    cores = O.use_all_cores()
    Z, xyz, ext = workloads.water_box_3k()
    for _ in range(min(args.warmup, 1)):
        O.solve(Z, xyz, 0.0, external=ext)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.solve(Z, xyz, 0.0, external=ext)
    wall = time.perf_counter() - t0
    value = wall / args.steps
    emit({"impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False,
          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "B3k: 3,000-atom water cluster + uniform field + 10,000 embedding point charges"},
          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": "every step is a full solve"},
          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


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
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "sm_mhz_min": min(sm) if sm else None,
                "power_w_max": max(power) if power else None}


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


def hbm_peak_gbs():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs")
    except Exception:
        return None


def int8_peak_tops():
    """Dense int8 tensor peak for the tcgen05 slice-product kernel, in 10^12 operations per second.  MEASURED_PEAKS.json
    holds the bf16 figure; int8 runs at twice that rate nominally.  The larger of that and the int8 issue rate this
    repository measured itself (tools/tc05_probe: back-to-back M128 N256 K32 tcgen05.mma.kind::i8 on every SM) is used, so
    the fraction is never flattered."""
    probe = (profile_record("int8_issue_rate") or {}).get("tops")
    try:
        bf16 = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops"])
        src = f"2 x MEASURED_PEAKS.json bf16_tflops (burst) = {2 * bf16:.0f}"
    except Exception:
        bf16, src = 1590.0, "2 x 1590 (fallback bf16 figure of B200_PROFILING.md)"
    peak = 2.0 * bf16
    if probe and probe > peak:
        peak, src = float(probe), src + f"; the measured tcgen05 int8 issue rate {probe:.0f} is higher and is used (profiles/r2_tcgen05_int8_probe.json)"
    return peak, src


def profile_record(key):
    """Facts that only an ncu capture can give (DRAM traffic, pipe utilisation) live in profiles/r2_ncu_facts.json,
    written from the committed ncu summaries -- never typed into this file."""
    try:
        return json.loads((ROOT / "profiles" / "r2_ncu_facts.json").read_text()).get(key)
    except Exception:
        return None


def md_throughput(ctx, rank, world, device, barrier, n_frames, steps, warmup):
    """Batched MD frames (BASELINE config 4): the trajectory is sharded round-robin over the ranks, every rank
    solves its frames with cheq_solve_batch (no data-path collective).  Device-timed with CUDA events on the
    engine's stream; returns seconds per pass over the whole trajectory (max over ranks) and details."""
    import torch
    import torch.distributed as dist

    from cheq_b200 import QEqSolver, SolverOptions, get_default_parameters, workloads
    from cheq_b200.parallel import shard_indices
    P = get_default_parameters()
    solver = QEqSolver(P, SolverOptions(), ctx)
    Zm, base = workloads.organic_chain(1000)
    mine = shard_indices(n_frames, rank, world)
    frames = workloads.md_frames_subset(base, mine)
    stream = torch.cuda.ExternalStream(ctx._lib.cheq_ctx_stream(ctx.handle), device=device)
    out = None
    for _ in range(max(1, warmup)):
        out = solver.solve_batch(Zm, frames[: max(64, len(frames) // 8)], 0.0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(stream)
    for _ in range(steps):
        out = solver.solve_batch(Zm, frames, 0.0)
        launches += ctx.stats()["kernel_launches"]
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    q, mu, its, status = out
    st = ctx.stats()
    ok = torch.tensor([float((status == 0).all()), float(its.min()), -float(its.max())], device=device)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return {"seconds_per_pass": float(ms.item()) / steps / 1e3, "frames": n_frames, "atoms": len(Zm),
            "all_converged": bool(ok[0].item() > 0.5), "iterations_min_max": [int(ok[1].item()), int(-ok[2].item())],
            "launches": launches, "flops_factor_per_pass_rank0": st["flops_factor"],
            "h2d_bytes": int(frames.nbytes + 17 * len(Zm)), "d2h_bytes": int(q.nbytes + mu.nbytes + its.nbytes + status.nbytes)}


def run_md(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import cheq_b200
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = cheq_b200.Context(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    r = md_throughput(ctx, rank, world, device, barrier, args.frames, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    lt = torch.tensor([float(r["launches"])], device=device)
    if world > 1:
        dist.all_reduce(lt)
    if rank == 0:
        peak_burst, peak_sustained = measure_fp64_peak(torch, device)
        fps = r["frames"] / r["seconds_per_pass"]
        tf = r["flops_factor_per_pass_rank0"] * world / r["seconds_per_pass"] / 1e12
        emit({"metric": metric_name(args), "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": r["seconds_per_pass"] * 1e3, "higher_is_better": True,
              "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": f"MD: {r['frames']} frames x {r['atoms']}-atom organic chain (C/H/N/O), STO, "
                                     f"hydrogen SCF, frames sharded round-robin over {world} rank(s)",
                         "l2": "every pass streams 16 MB of matrix per frame x thousands of frames >> 126 MB L2",
                         "iterations_min_max": r["iterations_min_max"], "all_converged": r["all_converged"]},
              "roofline": {"bound": "tensor", "kernel": "whole batched solve (GEMM + diagonal tiles + vector kernels)",
                           "achieved": tf / world, "peak": peak_sustained, "unit": "TFLOP/s",
                           "frac": tf / world / peak_sustained if peak_sustained else None, "traffic": None,
                           "note": "LAPACK-count flops of the factorisation work performed / wall time, per GPU"},
              "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": r["h2d_bytes"] * world,
                      "d2h_bytes_per_step": r["d2h_bytes"] * world,
                      "note": "cheq_solve_batch takes host buffers: value IS the end-to-end number"},
              "gpu_launches": int(lt.item()), "clocks": clocks})
    if world > 1:
        dist.destroy_process_group()


def secondary_metrics(ctx, rank, world, device, barrier):
    """BASELINE configs 3 and 4 in brief, reported next to the headline: batched MD frames/s on a 2,048-frame slice per
    rank (device-timed; `--workload md` runs the stated 10,000) and the B3k solve with field + point charges (rank 0)."""
    import torch

    from cheq_b200 import QEqSolver, SolverOptions, _ffi, get_default_parameters, workloads
    P = get_default_parameters()
    out = {}
    r = md_throughput(ctx, rank, world, device, barrier, 2048 * world, 2, 1)
    out["md"] = {"workload": f"{r['frames']} frames x {r['atoms']}-atom organic chain (C/H/N/O), STO, hydrogen SCF, "
                             f"frames sharded round-robin over {world} rank(s); CUDA-event timed",
                 "frames_per_s": r["frames"] / r["seconds_per_pass"], "seconds": r["seconds_per_pass"],
                 "iterations_min_max": r["iterations_min_max"], "all_converged": r["all_converged"]}
    if rank == 0:
        import cheq_b200
        solver = QEqSolver(P, SolverOptions(), cheq_b200.Context(device.index) if world > 1 else ctx)
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
        for _ in range(5):
            res = solver.solve_arrays(x, chi, hard, rad, nq, 0.0, e)
        st = solver.context.stats()
        out["b3k"] = {"workload": "3,000-atom water cluster + uniform field + 10,000 embedding point charges",
                      "ms_per_solve": (time.perf_counter() - t0) / 5 * 1e3, "iterations": res.iterations,
                      "krylov_iterations": st["krylov_iterations"], "kernel_launches": st["kernel_launches"]}
    return out


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import cheq_b200
    from cheq_b200 import SolverOptions
    from cheq_b200.solver import make_options, raise_for_status

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    name, Z, xyz, chi, hard, rad, nq = workload(n_molecules(args))   # one system, identical on every rank
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

    # N > 1: the collective answer must equal a private single-GPU solve of the same system (rank 0), before timing
    dist_check = None
    if world > 1:
        step_device()
        q_group = d_q.cpu().numpy().copy()
        mu_group, it_group = pot.value, int(its.value)
        if rank == 0 and n <= 40000:
            solo = cheq_b200.Context(local_rank)
            d_q1 = torch.empty_like(d_q)
            p1, i1, dl1 = C.c_double(), C.c_uint32(), C.c_double()
            rc = lib.cheq_solve_device(solo.handle, n, d_xyz.data_ptr(), d_chi.data_ptr(), d_hard.data_ptr(),
                                       d_rad.data_ptr(), d_nq.data_ptr(), 0.0, C.byref(opt), None, d_q1.data_ptr(),
                                       C.byref(p1), C.byref(i1), C.byref(dl1))
            raise_for_status(solo, rc, opt.max_iterations, dl1.value)
            dqs = float(np.max(np.abs(d_q1.cpu().numpy() - q_group)))
            dist_check = {"vs": "private single-GPU engine on rank 0", "dq": dqs, "dmu": abs(p1.value - mu_group),
                          "iterations": [it_group, int(i1.value)]}
            assert dqs < 1e-10 and abs(p1.value - mu_group) < 1e-10 and it_group == int(i1.value), dist_check
            solo.close()
            del d_q1
            torch.cuda.empty_cache()

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
    stats_plain = ctx.stats()
    trace = ctx.scf_trace()
    # pass 2: the same K steps with every launch of the dominant kernel bracketed by CUDA events on the engine's
    # stream (two event records per launch cost a few % of the step, hence not in pass 1)
    lib.cheq_set_profiling(ctx.handle, 1)
    gemm_ms = gemm_fl = tc_ms = tc_fl = tc_ops = 0.0
    gemm_n = tc_n = 0
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(args.steps):
        step_device()
        st = ctx.stats()
        gemm_ms += st["gemm_ms"]
        gemm_fl += st["gemm_flops"]
        gemm_n += st["gemm_launches"]
        tc_ms += st["tc_gemm_ms"]
        tc_fl += st["tc_gemm_flops"]
        tc_ops += st["tc_gemm_ops"]
        tc_n += st["tc_gemm_launches"]
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
    sampler_e2e = ClockSampler(local_rank)
    if rank == 0:
        sampler_e2e.start()
    ms_e2e = timed(step_host, args.steps)
    clocks_e2e = sampler_e2e.stop() if rank == 0 else None
    stats_e2e = ctx.stats()
    assert abs(pot.value - mu_dev) < 1e-9 and np.max(np.abs(h_q.numpy() - q_dev)) < 1e-9

    secondary = None
    if not args.no_secondary and args.workload == "s20k":
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
    # the updates split into two kernels: deep ones as int8 slice products on tcgen05 (oz_gemm_kernel), the rest on DMMA
    dmma_ms, dmma_fl, dmma_n = gemm_ms - tc_ms, gemm_fl - tc_fl, gemm_n - tc_n
    achieved = dmma_fl / (dmma_ms * 1e-3) / 1e12 if dmma_ms > 0 else 0.0
    dmma_block = {"bound": "tensor", "kernel": "gemm_nt_kernel (FP64 DMMA.8x8x4)", "achieved": achieved,
                  "peak": peak_sustained, "unit": "TFLOP/s", "frac": achieved / peak_sustained if peak_sustained else None,
                  "traffic": None, "launches": dmma_n, "kernel_ms_per_step": dmma_ms / args.steps,
                  "share_of_step": dmma_ms / ms_profiled,
                  "largest_launch": profile_record("gemm_largest_launch") if (n == 20001 and tc_n == 0) else None,
                  "peak_source": f"cuBLAS DGEMM 8192^3 measured in this run: sustained {peak_sustained:.1f}, "
                                 f"burst {peak_burst:.1f} TFLOP/s (MEASURED_PEAKS.json has no fp64 entry)"}
    tc_block = None
    if tc_n > 0:
        int8_peak, int8_src = int8_peak_tops()
        tc_tops = tc_ops / (tc_ms * 1e-3) / 1e12
        tc_block = {"bound": "tensor",
                    "kernel": "oz_gemm_kernel (tcgen05.mma.kind::i8 slice products with TMEM accumulators: the FP64-equivalent "
                              "updates C -= A S B^T), timed together with its slicing passes",
                    "achieved": tc_tops, "peak": int8_peak, "unit": "TFLOP/s", "frac": tc_tops / int8_peak,
                    "achieved_is": "int8 tensor operations per second (2 x multiply-adds, in 10^12/s): algorithmic FP64 "
                                   "flops x slice products per entry (28 at 7 slices, 15 at 5)",
                    "fp64_equivalent_tflops": tc_fl / (tc_ms * 1e-3) / 1e12,
                    "fp64_equivalent_vs_dgemm_peak": tc_fl / (tc_ms * 1e-3) / 1e12 / peak_sustained if peak_sustained else None,
                    "traffic": (profile_record("oz_largest_launch") or {}).get("traffic") if n == 20001 else None,
                    "launches": tc_n, "kernel_ms_per_step": tc_ms / args.steps, "share_of_step": tc_ms / ms_profiled,
                    "largest_launch": profile_record("oz_largest_launch") if n == 20001 else None,
                    "peak_source": int8_src}
    hbm_peak = hbm_peak_gbs()
    asm_gbs = stats["bytes_assemble"] / (stats["ms_assemble"] * 1e-3) / 1e9 if stats["ms_assemble"] > 0 else 0.0
    asm_gbs_per_gpu = asm_gbs / world          # every rank assembles 1/world of the tile columns in that time
    kry_gbs = (stats_plain["krylov_bytes"] / (stats_plain["krylov_ms"] * 1e-3) / 1e9
               if stats_plain["krylov_ms"] > 0 else None)

    # CPU legs: parity against ONE real oracle solve, which is also the CPU baseline
    full = None
    if args.workload == "s20k":
        full = load_oracle_result(name, n)
        if full is None and world == 1 and not args.no_cpu_baseline:
            full, _ = oracle_full_solve(name, Z, xyz)
    parity = None
    if full is not None:
        deltas_gpu = np.array([t[0] for t in trace])
        dd = (float(np.max(np.abs(deltas_gpu - full["deltas"]))) if len(deltas_gpu) == len(full["deltas"]) else None)
        parity = {"dq": float(np.max(np.abs(q_dev - full["charges"]))), "dmu": abs(mu_dev - float(full["mu"])),
                  "iterations": iterations, "iterations_ref": int(full["iterations"]), "max_delta_trace_diff": dd,
                  "tolerance": {"dq": 1e-8, "dmu": 1e-8},
                  "reference": "oracle/oracle.py full solve of the same workload (bordered partial-pivot LU per "
                               "SCF iteration, implementation.rs:273-345, 557-603)"}
        parity["ok"] = bool(parity["dq"] <= 1e-8 and parity["dmu"] <= 1e-8 and iterations == int(full["iterations"]))

    modes = "".join(str(t[1]) for t in trace)
    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(name, n, iterations),
                   "parallelism": (f"one system over {world} GPUs: 1-D block-cyclic column panels ("
                                   f"{os.environ.get('CHEQ_TILES_PER_PANEL', '8 (>= 4 panels per rank) or 4')} x 128 columns), owner-computes "
                                   "assembly/updates, finished panels pushed to the peers over NVLink (copy engines, "
                                   "CUDA IPC mappings, flag + cuStreamWaitValue32) up to 4 GPUs or NCCL-broadcast "
                                   "beyond, panel- and tile-level look-ahead, replicated triangular solves "
                                   f"(CHEQ_DIST_P2P={os.environ.get('CHEQ_DIST_P2P', '1')})") if world > 1 else "single GPU",
                   "l2": "working set (3.3 GB matrix, 1.4 GB Schur block, 1.4 GB R) >> 126 MB L2; no flush needed",
                   "phases_ms": {k: stats_plain[k] for k in ("ms_host_prepare", "ms_h2d", "ms_assemble",
                                                             "ms_factor_once", "ms_scf", "ms_d2h")},
                   "scf_modes": modes + "  (0 = hydrogen block refactorised, 1 = frozen factor + GMRES)",
                   "gmres_steps": [t[2] for t in trace if t[1] == 1],
                   "krylov_ms": stats_plain["krylov_ms"],
                   "factor_tflops_overall": stats_plain["flops_factor"] / ((stats_plain["ms_factor_once"] + stats_plain["ms_scf"]) * 1e9),
                   "lu_fallback_frames": stats_plain["lu_fallback_frames"]},
        # the dominant kernel by time: the tcgen05 slice-product kernel when it ran and took longer than the DMMA one
        "roofline": dict(tc_block if (tc_block and tc_ms >= dmma_ms) else dmma_block,
                         measured_in=f"second pass of {args.steps} steps with per-launch CUDA events "
                                     f"({ms_profiled / args.steps:.1f} ms per step vs {ms_total / args.steps:.1f} uninstrumented)",
                         potrf_ms_per_step=stats["potrf_ms"], backsolve_ms_per_step=stats["backsolve_ms"],
                         all_updates_fp64_tflops=gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
                         note=("panel schedule: GEMMs of the main, panel and bulk streams run concurrently, so the "
                               "summed per-launch durations overstate the kernel's wall-clock share and understate "
                               "its rate" if (world > 1 or os.environ.get("CHEQ_FACTOR_MODE") == "1") else
                               "recursive schedule: one stream, launches do not overlap")),
        "roofline_other_gemm": (dmma_block if (tc_block and tc_ms >= dmma_ms) else tc_block),
        "roofline_matvec": {"bound": "hbm", "kernel": "tile_matvec_kernel (stale-factor SCF iterations: one pass over "
                                                      "the lower triangle of S0 + two over the upper triangle of R per step)",
                            "achieved": kry_gbs, "peak": hbm_peak, "unit": "GB/s",
                            "frac": (kry_gbs / hbm_peak) if (kry_gbs and hbm_peak) else None,
                            "traffic": profile_record("tile_matvec_traffic") if n == 20001 else None,
                            "note": "algorithmic bytes of all GMRES steps / device time of the Krylov iterations "
                                    "(includes the host-synchronous GMRES bookkeeping and, in the iteration that "
                                    "freezes the factor, the n_h^3/3-flop construction of R)"},
        "roofline_assembly": {"bound": "hbm", "kernel": "assemble_lower_kernel<STO>", "achieved": asm_gbs_per_gpu,
                              "peak": hbm_peak, "unit": "GB/s", "frac": asm_gbs_per_gpu / hbm_peak if hbm_peak else None,
                              "traffic": profile_record("assemble_traffic") if n == 20001 else None,
                              "note": "per GPU: algorithmic bytes 8 N (N + 1) / 2 / n_gpus over the assembly time",
                              "peak_source": "MEASURED_PEAKS.json hbm_gbs"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * (24 + 8 + 8 + 8 + 1)),
                "d2h_bytes_per_step": int(n * 8 + 64),
                "clocks": clocks_e2e,
                "phases_ms_last_step": {k: stats_e2e[k] for k in ("ms_total", "ms_host_prepare", "ms_h2d", "ms_assemble",
                                                                  "ms_factor_once", "ms_scf", "ms_d2h")}},
        "gpu_launches": int(launches_t.item()),
        "clocks": clocks,
        "parity": parity,
    }
    if dist_check is not None:
        line["dist_check"] = dist_check
    if secondary is not None:
        line["secondary"] = secondary
    if full is not None and os.environ.get("CHEQ_ORACLE_CACHE"):
        # development runs: the parity reference was pre-computed elsewhere (tools/dev_cache), so its CPU time says
        # nothing about this box
        line["cpu_baseline"] = None
        line["parity"]["reference"] += "; loaded from CHEQ_ORACLE_CACHE (pre-computed, CPU time not measured in this run)"
    elif full is not None:
        line["cpu_baseline"] = {"value": float(full["t_total"]), "unit": UNIT, "cores": int(full["cores"]), "kind": "port",
                                "sample": f"the whole workload: ONE complete CPU solve of {name} ({int(full['iterations'])} SCF "
                                          f"iterations, assembly {float(full['t_asm']):.1f} s, median LU "
                                          f"{float(np.median(full['t_lu'])):.2f} s), measured on this box and cached under "
                                          "/tmp; cheq-equivalent CPU path (restated; faer/sto-ns not available offline)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_b3k(args, rank, local_rank, world):
    """BASELINE config 3 on one GPU: seconds per solve through the public API (host buffers)."""
    if rank != 0:
        return
    import torch

    import cheq_b200
    from cheq_b200 import QEqSolver, SolverOptions, _ffi, get_default_parameters, workloads
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    ctx = cheq_b200.Context(local_rank)
    P = get_default_parameters()
    solver = QEqSolver(P, SolverOptions(), ctx)
    Z, xyz, ext = workloads.water_box_3k()
    chi, hard, rad, nq = workloads.gather_parameters(Z, P)
    _, _, prad, pnq = workloads.gather_parameters(ext["Z"], P)
    e = _ffi.External()
    pxyz, pq = np.ascontiguousarray(ext["xyz"]), np.ascontiguousarray(ext["q"])
    e.num_point_charges = len(pq)
    e.xyz, e.charge = pxyz.ctypes.data_as(C.c_void_p), pq.ctypes.data_as(C.c_void_p)
    e.radius, e.n = prad.ctypes.data_as(C.c_void_p), pnq.ctypes.data_as(C.c_void_p)
    e.field[0], e.field[1], e.field[2] = ext["field"]
    x = np.ascontiguousarray(xyz)
    stream = torch.cuda.ExternalStream(ctx._lib.cheq_ctx_stream(ctx.handle), device=device)
    for _ in range(args.warmup):
        res = solver.solve_arrays(x, chi, hard, rad, nq, 0.0, e)
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(stream)
    for _ in range(args.steps):
        res = solver.solve_arrays(x, chi, hard, rad, nq, 0.0, e)
        launches += ctx.stats()["kernel_launches"]
    e1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    value = e0.elapsed_time(e1) / args.steps / 1e3
    st = ctx.stats()
    peak_burst, peak_sustained = measure_fp64_peak(torch, device)
    tf = st["flops_factor"] / value / 1e12
    emit({"metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "B3k: 3,000-atom water cluster + uniform field + 10,000 embedding point charges, STO, "
                                 f"hydrogen SCF ({res.iterations} iterations)",
                     "l2": "72 MB matrix fits the 126 MB L2: steps are back to back, L2-warm like the production loop",
                     "scf_modes": "".join(str(t[1]) for t in ctx.scf_trace())},
          "roofline": {"bound": "tensor", "kernel": "whole solve", "achieved": tf, "peak": peak_sustained,
                       "unit": "TFLOP/s", "frac": tf / peak_sustained if peak_sustained else None, "traffic": None},
          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(len(Z) * 49 + len(pq) * 41),
                  "d2h_bytes_per_step": int(len(Z) * 8 + 64), "note": "solve_arrays takes host buffers"},
          "gpu_launches": launches, "clocks": clocks})


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
    if args.workload == "md":
        run_md(args, rank, local_rank, world)
    elif args.workload == "b3k":
        run_b3k(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
