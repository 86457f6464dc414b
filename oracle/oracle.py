"""CPU oracle for the QEq hot path -- TEST INFRASTRUCTURE ONLY.

Restates caltechmsc/cheq's `QEqSolver::solve` / `solve_in_field`
(/root/reference/src/solver/implementation.rs:174-188, 239-266) on the CPU so that the B200 path can be
checked against it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; the product package (cheq_b200) never does.

Pair integrals, the external potential and the bordered-matrix assembly are plain C (oracle/cheq_oracle.c,
OpenMP row-parallel like the reference's rayon loop).  The partial-pivot LU of
`work_matrix.partial_piv_lu().solve(&rhs)` (implementation.rs:575, un-vendored crate faer ^0.23.2) is
LAPACK dgetrf/dgetrs through scipy; the SCF loop below follows implementation.rs:273-345 and 557-603
statement for statement.

Parity status: GTO pinned to 1e-8 eV (reference KATs, gto.rs:182-197); STO pinned to the 4 printed
decimals of BENCHMARKS.md (134/134 values reproduced, tests/test_oracle_golden.py); below that the STO
oracle is pinned to 40-digit mpmath values of the integral's definition, not to sto-ns itself
("parity unpinned" with respect to the crate: no Rust toolchain here).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import tomllib
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

H_CHARGE_DEPENDENCE_FACTOR = 0.93475415965  # implementation.rs:24
BOHR_TO_ANGSTROM = 0.529177210903  # constants.rs:13
HARTREE_TO_EV = 27.211386245988  # constants.rs:22

BASIS_GTO, BASIS_STO = 0, 1


def build(force: bool = False) -> Path:
    so = _HERE / "libcheq_oracle.so"
    src = _HERE / "cheq_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libcheq_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        d, u, i64, p = ctypes.c_double, ctypes.c_uint, ctypes.c_int64, ctypes.c_void_p
        _LIB.cheq_oracle_slater_exponent.restype = d
        _LIB.cheq_oracle_slater_exponent.argtypes = [u, d, d]
        _LIB.cheq_oracle_gto_integral.restype = d
        _LIB.cheq_oracle_gto_integral.argtypes = [d, u, d, u, d, d]
        _LIB.cheq_oracle_sto_integral.restype = d
        _LIB.cheq_oracle_sto_integral.argtypes = [d, u, d, u, d, d]
        _LIB.cheq_oracle_sto_coulomb.restype = d
        _LIB.cheq_oracle_sto_coulomb.argtypes = [d, u, d, u, d]
        _LIB.cheq_oracle_external_potential.restype = None
        _LIB.cheq_oracle_external_potential.argtypes = [i64, p, p, p, i64, p, p, p, p, p, d,
                                                        ctypes.c_int, p]
        _LIB.cheq_oracle_build_system.restype = i64
        _LIB.cheq_oracle_build_system.argtypes = [i64, p, p, p, p, p, d, p, d, ctypes.c_int, p, p, p]
        _LIB.cheq_oracle_num_threads.restype = ctypes.c_int
        _LIB.cheq_oracle_set_threads.restype = None
        _LIB.cheq_oracle_set_threads.argtypes = [ctypes.c_int]
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---- parameters (params.rs:21-45; resources/qeq.data.toml) ---------------------------------------
def default_parameters() -> dict[int, tuple[float, float, float, int]]:
    """Z -> (chi, hardness, radius, n).  Reads the repo's own copy of the element table."""
    path = _HERE.parent / "cheq_b200" / "data" / "qeq_params.toml"
    with open(path, "rb") as f:
        el = tomllib.load(f)["elements"]
    return {int(k): (v["chi"], v["j"], v["radius"], int(v["n"])) for k, v in el.items()}


@dataclass
class Options:  # options.rs:53-100
    tolerance: float = 1.0e-6
    max_iterations: int = 2000
    lambda_scale: float = 0.5
    hydrogen_scf: bool = True
    basis: int = BASIS_STO
    damping: tuple = ("auto", 0.4)  # ("none",) | ("fixed", d) | ("auto", initial)


class OracleError(Exception):
    def __init__(self, kind, **kw):
        super().__init__(f"{kind} {kw}")
        self.kind = kind
        self.info = kw


def gto_integral(R, n1, r1, n2, r2, lam):
    return lib().cheq_oracle_gto_integral(R, n1, r1, n2, r2, lam)


def sto_integral(R, n1, r1, n2, r2, lam):
    return lib().cheq_oracle_sto_integral(R, n1, r1, n2, r2, lam)


def sto_coulomb(R, n1, z1, n2, z2):
    return lib().cheq_oracle_sto_coulomb(R, n1, z1, n2, z2)


def gather(Z, params):
    """fetch_element_data (implementation.rs:348-362): first missing Z -> ParameterNotFound."""
    chi = np.empty(len(Z)); hard = np.empty(len(Z)); rad = np.empty(len(Z))
    nq = np.empty(len(Z), dtype=np.uint8)
    for i, z in enumerate(Z):
        if int(z) not in params:
            raise OracleError("ParameterNotFound", z=int(z))
        chi[i], hard[i], rad[i], nq[i] = params[int(z)]
    return chi, hard, rad, nq


def external_potential(xyz, rad, nq, pc_xyz, pc_q, pc_rad, pc_nq, field, opt: Options):
    """compute_external_potential (implementation.rs:369-447)."""
    n, m = len(rad), len(pc_q)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(n, 3)
    pc_xyz = np.ascontiguousarray(pc_xyz, dtype=np.float64).reshape(m, 3)
    pc_q = np.ascontiguousarray(pc_q, dtype=np.float64)
    pc_rad = np.ascontiguousarray(pc_rad, dtype=np.float64)
    pc_nq = np.ascontiguousarray(pc_nq, dtype=np.uint8)
    field = np.ascontiguousarray(field, dtype=np.float64)
    v = np.empty(n)
    lib().cheq_oracle_external_potential(n, _ptr(xyz), _ptr(rad), _ptr(nq), m, _ptr(pc_xyz), _ptr(pc_q),
                                         _ptr(pc_rad), _ptr(pc_nq), _ptr(field), opt.lambda_scale,
                                         opt.basis, _ptr(v))
    return v


def build_system(xyz, chi, hard, rad, nq, total_charge, v_ext, opt: Options):
    """build_invariant_system (implementation.rs:454-554) -> (matrix (n+1)^2 col-major, rhs, hydrogen idx)."""
    n = len(chi)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(n, 3)
    M = np.empty((n + 1, n + 1), order="F")
    rhs = np.empty(n + 1)
    hidx = np.empty(n, dtype=np.int64)
    vp = _ptr(v_ext) if v_ext is not None else None
    nh = lib().cheq_oracle_build_system(n, _ptr(xyz), _ptr(chi), _ptr(hard), _ptr(rad), _ptr(nq),
                                        float(total_charge), vp, opt.lambda_scale, opt.basis, _ptr(M),
                                        _ptr(rhs), _ptr(hidx))
    return M, rhs, hidx[:nh].copy()


@dataclass
class Result:  # types.rs:60-76
    charges: np.ndarray
    equilibrated_potential: float
    iterations: int
    deltas: list


def _lu_solve(work, rhs):
    """work_matrix.partial_piv_lu().solve(&rhs) (implementation.rs:575).  `work` is left intact, like faer."""
    from scipy.linalg import lu_factor, lu_solve
    lu, piv = lu_factor(work, overwrite_a=False, check_finite=False)
    x = lu_solve((lu, piv), rhs, check_finite=False)
    if not np.all(np.isfinite(x)):
        raise OracleError("LinalgError", msg="Linear system solver panicked. Matrix might be singular.")
    return x


def run_scf(n, base, rhs, hidx, hard, opt: Options, lu_solve=_lu_solve) -> Result:
    """run_scf_iterations + run_single_solve (implementation.rs:273-345, 557-603)."""
    charges = np.zeros(n)
    hydrogen_scf = opt.hydrogen_scf and len(hidx) > 0
    work = base.copy(order="F")
    if not hydrogen_scf:
        sol = lu_solve(work, rhs)
        charges = (1.0 - 1.0) * charges + 1.0 * sol[:n]
        return Result(charges, float(sol[n]), 1, [float(np.max(np.abs(sol[:n])))])
    delta = 0.0
    prev = np.finfo(np.float64).max
    kind = opt.damping[0]
    damping = 1.0 if kind == "none" else float(opt.damping[1])
    deltas = []
    for iteration in range(1, opt.max_iterations + 1):
        if iteration > 1:
            work[...] = base
        q_cl = np.clip(charges[hidx], -0.95, 0.95)
        work[hidx, hidx] = hard[hidx] * (1.0 + q_cl * H_CHARGE_DEPENDENCE_FACTOR)
        sol = lu_solve(work, rhs)
        new = sol[:n]
        delta = float(np.max(np.abs(new - charges)))
        charges = (1.0 - damping) * charges + damping * new
        mu = float(sol[n])
        deltas.append(delta)
        if delta < opt.tolerance:
            return Result(charges, mu, iteration, deltas)
        if kind == "auto":
            if delta > prev:
                damping *= 0.5
            elif delta < prev * 0.9:
                damping = min(damping * 1.1, 1.0)
            if damping < 0.001:
                damping = 0.001
        prev = delta
    raise OracleError("NotConverged", max_iterations=opt.max_iterations, delta=delta)


def solve(Z, xyz, total_charge=0.0, opt: Options | None = None, params=None, external=None) -> Result:
    """QEqSolver::solve / solve_in_field.  external = dict(Z=[...], xyz=[[..]], q=[...], field=[ex,ey,ez])."""
    opt = opt or Options()
    params = params or default_parameters()
    n = len(Z)
    ext_empty = external is None or (len(external.get("Z", [])) == 0 and
                                     all(float(c) == 0.0 for c in external.get("field", [0, 0, 0])))
    if n == 0:
        raise OracleError("NoAtoms")
    chi, hard, rad, nq = gather(Z, params)
    v_ext = None
    if not ext_empty:
        pz = external.get("Z", [])
        _, _, prad, pnq = gather(pz, params)
        v_ext = external_potential(xyz, rad, nq, external.get("xyz", np.zeros((0, 3))),
                                   external.get("q", []), prad, pnq,
                                   external.get("field", [0.0, 0.0, 0.0]), opt)
    base, rhs, hidx = build_system(xyz, chi, hard, rad, nq, total_charge, v_ext, opt)
    return run_scf(n, base, rhs, hidx, hard, opt)


def num_threads() -> int:
    return int(lib().cheq_oracle_num_threads())


def use_all_cores() -> int:
    """Give both halves of the CPU path every core this process may run on: the OpenMP assembly loops (the
    reference's rayon loop) and the LAPACK LU behind scipy (the reference's faer).  Needed under launchers that
    export OMP_NUM_THREADS=1 (torchrun).  Returns the core count."""
    import os
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().cheq_oracle_set_threads(cores)
    try:
        from threadpoolctl import threadpool_limits
        global _BLAS_LIMIT
        _BLAS_LIMIT = threadpool_limits(limits=cores)      # stays in force for the life of the process
    except Exception:
        pass
    return cores


_BLAS_LIMIT = None
