#!/usr/bin/env python3
"""Matrix-free path on a water cluster of the given size: time, operator applications, pair rate; compared with the
dense path when the latter fits.  Run under torchrun to shard the operator's rows over the ranks."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cheq_b200
from cheq_b200 import QEqSolver, SolverOptions, get_default_parameters, workloads

mol = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
dense_too = "--no-dense" not in sys.argv
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch, torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = cheq_b200.Context(local)
if world > 1:
    from cheq_b200 import distributed
    distributed.join(ctx)
P = get_default_parameters()
Z, xyz = workloads.water_cluster(mol)
chi, hard, rad, nq = workloads.gather_parameters(Z, P)
x = np.ascontiguousarray(xyz)
s = QEqSolver(P, SolverOptions(), ctx)
t0 = time.perf_counter()
it = s.solve_arrays(x, chi, hard, rad, nq, 0.0, None, iterative=True)
dt = time.perf_counter() - t0
st = ctx.stats(); tr = ctx.scf_trace()
if rank == 0:
    print(f"matrix-free: n={len(Z)} world={world} iterations={it.iterations} mu={it.equilibrated_potential:.10f} "
          f"time={dt:.2f} s  operator applications={st['matfree_matvecs']}  pairs/s={st['matfree_pairs'] / dt:.3e} "
          f"(whole solve)  GMRES steps per SCF iteration {[t[2] for t in tr]}  sum(q)={sum(it.charges):.2e}", flush=True)
if dense_too:
    t0 = time.perf_counter()
    d = s.solve_arrays(x, chi, hard, rad, nq, 0.0)
    dt2 = time.perf_counter() - t0
    if rank == 0:
        dq = float(np.max(np.abs(np.array(d.charges) - np.array(it.charges))))
        print(f"dense: iterations={d.iterations} time={dt2:.2f} s (first call, includes allocation); matrix-free vs dense: "
              f"max|dq|={dq:.2e} |dmu|={abs(d.equilibrated_potential - it.equilibrated_potential):.2e}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
