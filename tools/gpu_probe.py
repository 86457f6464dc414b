#!/usr/bin/env python3
"""Timing probe run on the GPU box: per-phase times of single solves at a few sizes."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheq_b200
from cheq_b200 import QEqSolver, SolverOptions, get_default_parameters, workloads


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1000, 6667]
    P = get_default_parameters()
    ctx = cheq_b200.default_context(0)
    s = QEqSolver(P, SolverOptions(), ctx)
    ctx._lib.cheq_set_profiling(ctx.handle, 1)
    for nm in sizes:
        Z, xyz = workloads.water_cluster(nm)
        chi, hard, rad, nq = workloads.gather_parameters(Z, P)
        xyz = np.ascontiguousarray(xyz)
        for rep in range(2):
            t0 = time.time()
            r = s.solve_arrays(xyz, chi, hard, rad, nq, 0.0)
            dt = time.time() - t0
            st = ctx.stats()
            print(f"n={len(Z)} rep={rep} wall={dt*1e3:.1f} ms its={r.iterations} mu={r.equilibrated_potential:.9f} "
                  f"sumq={sum(r.charges):.2e} | host {st['ms_host_prepare']:.1f} h2d {st['ms_h2d']:.2f} "
                  f"asm {st['ms_assemble']:.2f} once {st['ms_factor_once']:.1f} scf {st['ms_scf']:.1f} "
                  f"d2h {st['ms_d2h']:.2f} launches {st['kernel_launches']} "
                  f"TF/s {st['flops_factor']/((st['ms_factor_once']+st['ms_scf'])*1e9):.2f} | gemm {st['gemm_ms']:.1f} ms "
                  f"({st['gemm_flops']/max(st['gemm_ms'],1e-9)/1e9:.1f} TF/s, {st['gemm_launches']} launches) potrf {st['potrf_ms']:.1f} ms "
                  f"backsolve {st['backsolve_ms']:.1f} ms", flush=True)


if __name__ == "__main__":
    main()
