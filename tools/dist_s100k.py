#!/usr/bin/env python3
"""BASELINE config 5: the 100,000-atom solvated system (STO) as ONE dense QEq solve over the GPUs of a node.

  python -m torch.distributed.run --nproc-per-node 8 tools/dist_s100k.py [--atoms-scale 1.0]

Checks (size-independent properties; the CPU oracle cannot factor a 100k system): charge neutrality, identical
results on every rank, stationarity chi_i + sum_j A_ij q_j + mu = 0 on sampled rows of the single-factorisation
solve (rows recomputed with the CPU oracle's integral), then times the full hydrogen-SCF solve.  Prints one JSON
line on rank 0."""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
BOHR, EV = 0.529177210903, 27.211386245988


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--solute", type=int, default=10000)
    ap.add_argument("--waters", type=int, default=30000)
    ap.add_argument("--rows", type=int, default=8)
    ap.add_argument("--skip-scf", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import cheq_b200
    from cheq_b200 import QEqSolver, SolverOptions, distributed, get_default_parameters, workloads

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = cheq_b200.Context(local)
    if world > 1:
        distributed.join(ctx)
    P = get_default_parameters()
    Z, xyz = workloads.solvated_100k(args.solute, args.waters)
    n = len(Z)
    xyz = np.ascontiguousarray(xyz)
    chi, hard, rad, nq = workloads.gather_parameters(Z, P)
    out = {"workload": f"S100k: {args.solute}-atom organic solute + {args.waters} waters = {n} atoms, STO",
           "n_gpus": world, "n_hydrogen": int((nq == 1).sum())}

    def agree(q, mu):
        if world == 1:
            return True
        t = torch.tensor(np.concatenate([q, [mu]]), device=device)
        ref = t.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([1.0 if torch.equal(t, ref) else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(ok.item() == 1.0)

    # 1. single factorisation (hydrogen_scf off): exact linear system -> stationarity on sampled rows
    s1 = QEqSolver(P, SolverOptions(hydrogen_scf=False), ctx)
    t0 = time.perf_counter()
    r1 = s1.solve_arrays(xyz, chi, hard, rad, nq, 0.0)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    r1 = s1.solve_arrays(xyz, chi, hard, rad, nq, 0.0)
    t_single = time.perf_counter() - t0
    st = ctx.stats()
    q1 = np.array(r1.charges)
    out["single_factorisation"] = {"seconds": t_single, "seconds_first_call": t_first, "sum_q": float(q1.sum()),
                                   "ranks_agree": agree(q1, r1.equilibrated_potential),
                                   "lu_fallback": int(st["lu_fallback_frames"]), "ms_assemble": st["ms_assemble"],
                                   "ms_factor": st["ms_factor_once"], "flops_rank0": st["flops_factor"]}
    if rank == 0 and args.rows > 0:
        from oracle import oracle as O
        rng = np.random.default_rng(2)
        worst = 0.0
        for i in rng.choice(n, size=args.rows, replace=False):
            d = np.linalg.norm(xyz - xyz[i], axis=1) / BOHR
            row = np.array([O.sto_integral(d[j], int(nq[i]), rad[i] / BOHR, int(nq[j]), rad[j] / BOHR, 0.5) * EV
                            if j != i else hard[i] for j in range(n)])
            worst = max(worst, abs(chi[i] + row @ q1 + r1.equilibrated_potential))
        out["single_factorisation"]["max_stationarity_residual_eV"] = worst
        out["single_factorisation"]["rows_checked"] = args.rows
    # 2. full hydrogen-SCF solve (default options), timed
    if not args.skip_scf:
        s2 = QEqSolver(P, SolverOptions(), ctx)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r2 = s2.solve_arrays(xyz, chi, hard, rad, nq, 0.0)
        torch.cuda.synchronize()
        t_scf = time.perf_counter() - t0
        tt = torch.tensor([t_scf], device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        st = ctx.stats()
        q2 = np.array(r2.charges)
        fl = torch.tensor([st["flops_factor"]], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(fl)
        out["scf"] = {"seconds": float(tt.item()), "iterations": r2.iterations, "sum_q": float(q2.sum()),
                      "mu": r2.equilibrated_potential, "ranks_agree": agree(q2, r2.equilibrated_potential),
                      "lu_fallback": int(st["lu_fallback_frames"]), "ms_assemble": st["ms_assemble"],
                      "ms_factor_once": st["ms_factor_once"], "ms_scf": st["ms_scf"],
                      "flops_all_ranks": float(fl.item()),
                      "tflops_per_gpu": float(fl.item()) / world / max(1e-9, (st["ms_factor_once"] + st["ms_scf"]) * 1e9),
                      "q_min_max": [float(q2.min()), float(q2.max())]}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
