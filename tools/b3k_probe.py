#!/usr/bin/env python3
"""BASELINE config 3 (3,000-atom water cluster + uniform field + 10,000 embedding point charges): phase times of
one solve and of the external-potential kernel alone (cheq_external_potential)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheq_b200
from cheq_b200 import QEqSolver, SolverOptions, _ffi, get_default_parameters, workloads

P = get_default_parameters()
ctx = cheq_b200.default_context(0)
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
s = QEqSolver(P, SolverOptions(), ctx)
for rep in range(3):
    t0 = time.perf_counter()
    r = s.solve_arrays(x, chi, hard, rad, nq, 0.0, e)
    dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"B3k rep={rep} wall={dt * 1e3:.2f} ms its={r.iterations} host {st['ms_host_prepare']:.2f} h2d {st['ms_h2d']:.3f} "
          f"vext+assemble {st['ms_assemble']:.3f} once {st['ms_factor_once']:.2f} scf {st['ms_scf']:.2f} launches {st['kernel_launches']}")
v = np.empty(len(Z))
lib = ctx._lib
best = 1e9
for rep in range(5):
    t0 = time.perf_counter()
    rc = lib.cheq_external_potential(ctx.handle, len(Z), x.ctypes.data_as(C.c_void_p), rad.ctypes.data_as(C.c_void_p),
                                     nq.ctypes.data_as(C.c_void_p), C.byref(e), 0.5, 1, v.ctypes.data_as(C.c_void_p))
    best = min(best, time.perf_counter() - t0)
assert rc == 0
print(f"cheq_external_potential: best wall {best * 1e3:.3f} ms for {len(Z)} x {len(pq)} = {len(Z) * len(pq):.2e} STO integrals "
      f"(includes H2D of the charges and the matrix assembly that shares the call)")
