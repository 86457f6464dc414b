#!/usr/bin/env python3
"""In-situ launch list of ONE steady-state S20k solve through CUPTI (torch.profiler, CUDA activity only): per kernel
the launch count, summed and average device time and the share of the solve.  Unlike the ncu launch list
(serialised, cold-cache, ~90 ms of tool time per launch) the kernels run back to back as in the benchmark, so the
shares are the real ones.

    python tools/cupti_launches.py [waters] > profiles/r2_cupti_launches_s20k.txt
"""
import collections
import os
import re
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cheq_b200  # noqa: E402
from cheq_b200 import QEqSolver, SolverOptions, get_default_parameters, workloads  # noqa: E402


def main():
    waters = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
    torch.cuda.init()
    Z, xyz = workloads.water_cluster(waters, workloads.SEED)
    chi, hard, rad, nq = workloads.gather_parameters(Z, get_default_parameters())
    ctx = cheq_b200.Context(0)
    s = QEqSolver(get_default_parameters(), SolverOptions(), ctx)
    xyz = np.ascontiguousarray(xyz)
    for _ in range(2):
        res = s.solve_arrays(xyz, chi, hard, rad, nq, 0.0)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        res = s.solve_arrays(xyz, chi, hard, rad, nq, 0.0)
        torch.cuda.synchronize()
    st = ctx.stats()
    agg = collections.defaultdict(lambda: [0, 0.0])
    t_min, t_max = None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "").replace("cheq::", "").replace("void ", ""))
        dur = float(getattr(ev, "device_time", 0.0) or getattr(ev, "cuda_time", 0.0) or 0.0)
        agg[name][0] += 1
        agg[name][1] += dur
        tr = ev.time_range
        t_min = tr.start if t_min is None else min(t_min, tr.start)
        t_max = tr.end if t_max is None else max(t_max, tr.end)
    tot = sum(v[1] for v in agg.values())
    span = (t_max - t_min) if t_min is not None else 0.0
    print(f"# CUPTI activity records (torch.profiler, CUDA only) of one steady-state solve: {len(Z)} atoms, "
          f"{res.iterations} SCF iterations, {st['ms_total']:.1f} ms by the engine's host clock under the profiler")
    print(f"# summed device time of all records {tot / 1e3:.1f} ms over a span of {span / 1e3:.1f} ms "
          "(kernels of one stream: the difference is idle time between launches; memcpy / memset records included)")
    print(f"{'kernel / activity':48s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:48]:48s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%")
    print(f"{'total':48s} {sum(v[0] for v in agg.values()):8d} {tot:12.1f}")


if __name__ == "__main__":
    main()
