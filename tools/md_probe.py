#!/usr/bin/env python3
"""Batched-trajectory probe on the GPU box: frames/s of cheq_solve_batch on the MD workload (1,000-atom organic)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheq_b200
from cheq_b200 import QEqSolver, SolverOptions, get_default_parameters, workloads


def main():
    frames = [int(a) for a in sys.argv[1:]] or [64, 512]
    P = get_default_parameters()
    ctx = cheq_b200.default_context(0)
    s = QEqSolver(P, SolverOptions(), ctx)
    Z, base = workloads.organic_chain(1000)
    print("atoms", len(Z), "hydrogens", sum(1 for z in Z if z == 1))
    for F in frames:
        xyz = workloads.md_frames(base, F)
        for rep in range(2):
            t0 = time.time()
            q, mu, its, status = s.solve_batch(Z, xyz, 0.0)
            dt = time.time() - t0
            st = ctx.stats()
            print(f"F={F} rep={rep} wall={dt*1e3:.1f} ms -> {F/dt:.1f} frames/s | its min/max {its.min()}/{its.max()} "
                  f"status ok {int((status == 0).sum())} | host {st['ms_host_prepare']:.1f} h2d {st['ms_h2d']:.2f} "
                  f"asm {st['ms_assemble']:.2f} once {st['ms_factor_once']:.1f} scf {st['ms_scf']:.1f} "
                  f"launches {st['kernel_launches']} TF/s {st['flops_factor']/((st['ms_factor_once']+st['ms_scf'])*1e9):.2f}",
                  flush=True)
    single = s.solve([cheq_b200.Atom(z, p.tolist()) for z, p in zip(Z, xyz[0])], 0.0)
    print("single vs batch frame 0: max|dq| =", float(np.max(np.abs(np.array(single.charges) - q[0]))), "its", single.iterations)


if __name__ == "__main__":
    main()
