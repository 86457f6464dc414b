#!/usr/bin/env python3
"""Forced pivoted-LU fallback timing (CHEQ_FORCE_LU=1): the cost of the cliff a pivot breakdown would cause.
Prints one JSON line per size: seconds per solve on the regular path, with the fallback forced (regular + LU), and
the difference = the LU path itself (one partial-pivot LU of the bordered matrix per SCF iteration)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cheq_b200  # noqa: E402
from cheq_b200 import QEqSolver, SolverOptions, get_default_parameters, workloads  # noqa: E402


def timed(ctx, Z, xyz, reps):
    chi, hard, rad, nq = workloads.gather_parameters(Z, get_default_parameters())
    s = QEqSolver(get_default_parameters(), SolverOptions(), ctx)
    res = s.solve_arrays(np.ascontiguousarray(xyz), chi, hard, rad, nq, 0.0)
    t0 = time.perf_counter()
    for _ in range(reps):
        res = s.solve_arrays(np.ascontiguousarray(xyz), chi, hard, rad, nq, 0.0)
    return (time.perf_counter() - t0) / reps, res


for waters, reps in ((1000, 3), (6667, 1)):
    Z, xyz = workloads.water_cluster(waters, workloads.SEED)
    ctx = cheq_b200.Context(0)
    t_reg, r0 = timed(ctx, Z, xyz, reps)
    ctx.close()
    os.environ["CHEQ_FORCE_LU"] = "1"
    ctx = cheq_b200.Context(0)
    del os.environ["CHEQ_FORCE_LU"]
    t_lu, r1 = timed(ctx, Z, xyz, reps)
    st = ctx.stats()
    ctx.close()
    n = len(Z)
    its = r1.iterations
    print(json.dumps({"atoms": n, "iterations": its, "s_regular": t_reg, "s_forced_fallback": t_lu, "s_lu_path": t_lu - t_reg,
                      "lu_tflops": its * (2.0 / 3.0) * (n + 1) ** 3 / max(t_lu - t_reg, 1e-9) / 1e12,
                      "lu_fallback_frames": st["lu_fallback_frames"],
                      "max_dq_vs_regular": float(np.max(np.abs(np.array(r1.charges) - np.array(r0.charges))))}), flush=True)
