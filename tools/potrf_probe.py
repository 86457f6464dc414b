#!/usr/bin/env python3
"""Cycle breakdown of the blocked diagonal-tile kernel: runs a Cholesky solve through the debug entry point and
prints the per-launch averages of the in-kernel clock64 counters."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheq_b200

ctx = cheq_b200.Context(0)
lib = ctx._lib
n = 2048
rng = np.random.default_rng(0)
B = rng.normal(size=(n, n))
A = np.asfortranarray(B @ B.T + n * np.eye(n))
b = np.asfortranarray(rng.normal(size=(n, 1)))
x = np.empty((n, 1), order="F")
out = (C.c_uint64 * 8)()
lib.cheq_dbg_potrf_cycles(ctx.handle, out, 1)
for rep in range(3):
    rc = lib.cheq_dbg_cholesky_solve(ctx.handle, n, A.ctypes.data_as(C.c_void_p), 1, b.ctypes.data_as(C.c_void_p),
                                     x.ctypes.data_as(C.c_void_p), None)
    assert rc == 0, ctx.last_error()
lib.cheq_dbg_potrf_cycles(ctx.handle, out, 1)
v = list(out)
k = max(1, v[5])
print("launches", v[5], "cycles per launch: load %.0f A %.0f B %.0f C %.0f store %.0f total %.0f" %
      (v[0] / k, v[1] / k, v[2] / k, v[3] / k, v[4] / k, sum(v[:5]) / k))
print("residual", float(np.max(np.abs(A @ x - b))))
tf = C.c_double()
lib.cheq_dbg_dfma_rate(ctx.handle, C.byref(tf))
print("DFMA rate TF/s", tf.value, "|", ctx.last_error())
