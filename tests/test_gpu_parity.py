"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.  Needs a B200.

Tolerances (BASELINE.json north_star): 1e-8 e per atomic charge, 1e-8 eV in the equilibrated potential,
identical SCF iteration counts; pair integrals and matrix entries are held to 1e-13 relative.
"""
import ctypes as C
import math

import mpmath as mp
import numpy as np
import pytest

import cheq_b200
from cheq_b200 import (Atom, BasisType, DampingStrategy, ExternalPotential, PointCharge, QEqSolver, SolverOptions,
                       _ffi, get_default_parameters, workloads)
from conftest import rg_cases

pytestmark = pytest.mark.gpu

BOHR = 0.529177210903
EV = 27.211386245988
Q_TOL = 1e-8
MU_TOL = 1e-8


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def gpu_pair_integrals(ctx, dist_A, n1, r1, n2, r2, lam, basis):
    dist_A = np.ascontiguousarray(dist_A, dtype=np.float64)
    n1 = np.ascontiguousarray(n1, dtype=np.uint8)
    n2 = np.ascontiguousarray(n2, dtype=np.uint8)
    r1 = np.ascontiguousarray(r1, dtype=np.float64)
    r2 = np.ascontiguousarray(r2, dtype=np.float64)
    out = np.empty(len(dist_A))
    rc = ctx._lib.cheq_pair_integrals(ctx.handle, len(dist_A), _ptr(dist_A), _ptr(n1), _ptr(r1), _ptr(n2), _ptr(r2),
                                      lam, basis, _ptr(out))
    assert rc == 0, ctx.last_error()
    return out


def solver(ctx, **kw):
    return QEqSolver(get_default_parameters(), SolverOptions(**kw), ctx)


def atoms_of(Z, xyz):
    return [Atom(int(z), [float(c) for c in p]) for z, p in zip(Z, xyz)]


def oracle_opts(oracle, o: SolverOptions):
    return oracle.Options(tolerance=o.tolerance, max_iterations=o.max_iterations, lambda_scale=o.lambda_scale,
                          hydrogen_scf=o.hydrogen_scf, basis=o.basis_type.value,
                          damping=(o.damping.kind, o.damping.value))


def assert_parity(res, ref):
    assert res.iterations == ref.iterations, (res.iterations, ref.iterations)
    dq = float(np.max(np.abs(np.array(res.charges) - ref.charges)))
    dmu = abs(res.equilibrated_potential - ref.equilibrated_potential)
    assert dq <= Q_TOL, dq
    assert dmu <= MU_TOL, dmu
    return dq, dmu


# ---- pair integrals -------------------------------------------------------------------------------------------
def test_pair_integrals_sto_vs_definition(gpu_ctx, sto_truth):
    rows = [r for r in sto_truth["rows"] if r["z1"] != 0]      # element pairs of the default table
    P = get_default_parameters().elements
    d = np.array([float.fromhex(r["R_bohr"]) * BOHR for r in rows])
    n1 = np.array([r["n1"] for r in rows]); n2 = np.array([r["n2"] for r in rows])
    r1 = np.array([P[r["z1"]].radius for r in rows]); r2 = np.array([P[r["z2"]].radius for r in rows])
    out = gpu_pair_integrals(gpu_ctx, d, n1, r1, n2, r2, sto_truth["lambda"], 1)
    for r, v in zip(rows, out):
        t = mp.mpf(r["J_hartree"]) * mp.mpf(EV)
        R = float.fromhex(r["R_bohr"])
        rel = float(abs((mp.mpf(float(v)) - t) / t))
        assert rel < (1e-13 if R >= 0.5 else 1e-12), (r, v, rel)


def test_pair_integrals_vs_oracle(gpu_ctx, oracle):
    P = get_default_parameters().elements
    rng = np.random.default_rng(7)
    zs = [1, 2, 3, 6, 7, 8, 9, 11, 14, 15, 17, 19, 24, 25, 26, 29, 35, 37, 53, 55, 5, 23, 27, 42, 79, 92]
    pairs = [(a, b) for a in zs for b in zs]
    idx = rng.choice(len(pairs), size=400)
    z1 = np.array([pairs[i][0] for i in idx]); z2 = np.array([pairs[i][1] for i in idx])
    d = rng.uniform(0.5, 14.0, size=len(idx))
    d[:20] = rng.uniform(20.0, 200.0, size=20)
    n1 = np.array([P[z].principal_quantum_number for z in z1]); n2 = np.array([P[z].principal_quantum_number for z in z2])
    r1 = np.array([P[z].radius for z in z1]); r2 = np.array([P[z].radius for z in z2])
    for basis, lam in ((1, 0.5), (0, 0.5), (1, 0.4913), (0, 0.4913)):
        out = gpu_pair_integrals(gpu_ctx, d, n1, r1, n2, r2, lam, basis)
        f = oracle.sto_integral if basis == 1 else oracle.gto_integral
        ref = np.array([f(d[k] / BOHR, int(n1[k]), r1[k] / BOHR, int(n2[k]), r2[k] / BOHR, lam) * EV
                        for k in range(len(d))])
        rel = np.max(np.abs(out - ref) / np.abs(ref))
        assert rel < 1e-13, (basis, lam, rel)
    # GTO known answers (gto.rs:182-197)
    out = gpu_pair_integrals(gpu_ctx, [0.74, 1.128], [1, 2], [0.371, 0.759], [1, 2], [0.371, 0.669], 0.4913, 0)
    assert abs(out[0] - 14.02504752) < 1e-8 and abs(out[1] - 5.837573337) < 1e-8
    # R -> 0 limits are finite and continuous (sto.rs:84-94, gto.rs:154-166)
    for basis in (0, 1):
        o = gpu_pair_integrals(gpu_ctx, [0.0, 1e-10], [1, 1], [0.371, 0.371], [1, 1], [0.371, 0.371], 0.5, basis)
        assert np.all(np.isfinite(o)) and abs(o[0] - o[1]) < 1e-6


# ---- dense kernels ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 384, 256), (640, 128, 1152)])
def test_gemm_nt_dmma(gpu_ctx, m, n, k):
    rng = np.random.default_rng(m + n + k)
    A = np.asfortranarray(rng.normal(size=(m, k)))
    B = np.asfortranarray(rng.normal(size=(n, k)))
    C0 = np.asfortranarray(rng.normal(size=(m, n)))
    for sub in (1, 0):
        Cg = C0.copy(order="F")
        rc = gpu_ctx._lib.cheq_dbg_gemm_nt(gpu_ctx.handle, m, n, k, _ptr(A), _ptr(B), _ptr(Cg), sub)
        assert rc == 0, gpu_ctx.last_error()
        ref = C0 - A @ B.T if sub else A @ B.T
        assert np.max(np.abs(Cg - ref)) < 1e-11 * math.sqrt(k)


@pytest.mark.parametrize("m,n,k,slices,lower", [(128, 128, 128, 7, 0), (384, 256, 640, 4, 0), (384, 256, 640, 5, 0),
                                                 (384, 256, 640, 6, 0), (384, 256, 640, 7, 0), (384, 256, 640, 8, 0),
                                                 (640, 384, 1152, 7, 1), (2560, 2560, 2048, 7, 1), (256, 128, 8192, 7, 0)])
def test_gemm_ozaki_tcgen05(gpu_ctx, m, n, k, slices, lower):
    """C -= A S B^T as int8 slice products on tcgen05.mma.kind::i8 (kernels_ozaki.cu): the error of an entry is
    bounded relative to rowmax(A) * rowmax(B) * K by the slice count (54 bits at 7 slices), whatever the dynamic range
    of the rows; pivot signs are folded into B; with `lower` the tiles above the diagonal are not touched."""
    rng = np.random.default_rng(m + n + k + slices)
    A = rng.normal(size=(m, k)) * np.exp(2.0 * rng.normal(size=(m, 1))) * np.exp2(-rng.integers(0, 8, size=(m, k)))
    B = rng.normal(size=(n, k)) * np.exp(2.0 * rng.normal(size=(n, 1)))
    A[1, :] = 0.0                      # an all-zero row
    A[2, 3] = 1e12                     # one huge entry: the rest of that row loses relative, not absolute, accuracy
    A, B = np.asfortranarray(A), np.asfortranarray(B)
    C0 = np.asfortranarray(rng.normal(size=(m, n)))
    sgn = np.where(rng.random(k) < 0.3, -1.0, 1.0)
    scale = np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=1)[None, :] * k
    small = m * n * k <= 1 << 27
    prod = ((A * sgn).astype(np.longdouble) @ B.T.astype(np.longdouble)) if small else (A * sgn) @ B.T
    for sub in (1, 0):
        Cg = C0.copy(order="F")
        rc = gpu_ctx._lib.cheq_dbg_gemm_ozaki(gpu_ctx.handle, m, n, k, _ptr(A), _ptr(B), _ptr(Cg), sub, lower, slices,
                                              _ptr(sgn))
        assert rc == 0, gpu_ctx.last_error()
        ref = np.asarray(C0 - prod if sub else prod, dtype=np.longdouble)
        # beyond the slicing bound only the fp64 rounding of the recombination and of the update of C is allowed
        err = np.maximum(np.abs(Cg - ref) - 4e-16 * (np.abs(C0) + np.abs(np.asarray(prod, dtype=np.float64))), 0.0) / (scale + 1e-300)
        touched = np.ones((m, n), dtype=bool)
        if lower:
            i, j = np.indices((m, n))
            touched = (i // 128) * 128 + 127 >= (j // 64) * 64
            assert np.array_equal(Cg[~touched], C0[~touched])
        bound = max(2.0 ** -(8 * slices - 6), 1e-17 if small else 2e-15 / math.sqrt(k))
        assert float(err[touched].max()) <= bound, (float(err[touched].max()), bound)


@pytest.mark.parametrize("n", [5, 128, 200, 515, 1300])
def test_cholesky_solve(gpu_ctx, n):
    rng = np.random.default_rng(n)
    G = rng.normal(size=(n, n))
    A = np.asfortranarray(G @ G.T / n + np.eye(n) * 2.0)
    b = np.asfortranarray(rng.normal(size=(n, 2)))
    x = np.empty((n, 2), order="F")
    L = np.empty((n, n), order="F")
    rc = gpu_ctx._lib.cheq_dbg_cholesky_solve(gpu_ctx.handle, n, _ptr(A), 2, _ptr(b), _ptr(x), _ptr(L))
    assert rc == 0, gpu_ctx.last_error()
    assert np.max(np.abs(L - np.linalg.cholesky(A))) < 1e-12
    assert np.max(np.abs(x - np.linalg.solve(A, b))) < 1e-11
    # symmetric indefinite: the factorisation is A = L S L^T with S = diag(+-1) (no pivoting)
    A[0, 0] = -1.0
    A[n // 2, n // 2] -= 3.5
    rc = gpu_ctx._lib.cheq_dbg_cholesky_solve(gpu_ctx.handle, n, _ptr(A), 2, _ptr(b), _ptr(x), None)
    assert rc == 0, gpu_ctx.last_error()
    assert np.linalg.eigvalsh(A)[0] < 0
    assert np.max(np.abs(x - np.linalg.solve(A, b))) < 1e-10
    # a vanishing pivot is reported (the engine then switches to the pivoted LU)
    A[0, 0] = 0.0
    rc = gpu_ctx._lib.cheq_dbg_cholesky_solve(gpu_ctx.handle, n, _ptr(A), 2, _ptr(b), _ptr(x), None)
    assert rc == _ffi.LINALG


# ---- assembly -------------------------------------------------------------------------------------------------
def _assemble(ctx, Z, xyz, lam, basis):
    chi, hard, rad, nq = workloads.gather_parameters(Z, get_default_parameters())
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    n = len(Z)
    M = np.empty((n, n), order="F")
    rc = ctx._lib.cheq_assemble_matrix(ctx.handle, n, _ptr(xyz), _ptr(hard), _ptr(rad), _ptr(nq), lam, basis, _ptr(M))
    assert rc == 0, ctx.last_error()
    return M


@pytest.mark.parametrize("basis", [1, 0])
def test_assembled_matrix_vs_oracle(gpu_ctx, oracle, basis):
    cases = [workloads.water_cluster(27), workloads.organic_chain(300),
             ([3, 9, 11, 17, 55, 53, 26, 29, 1, 2, 79, 92, 14, 15],
              np.random.default_rng(3).uniform(0, 6, size=(14, 3)))]
    P = oracle.default_parameters()
    for Z, xyz in cases:
        M = _assemble(gpu_ctx, Z, xyz, 0.5, basis)
        chi, hard, rad, nq = oracle.gather(Z, P)
        ref, _, _ = oracle.build_system(xyz, chi, hard, rad, nq, 0.0, None, oracle.Options(basis=basis))
        n = len(Z)
        ref = ref[:n, :n]
        assert np.array_equal(M, M.T)
        assert np.max(np.abs(M - ref) / np.abs(ref)) < 1e-13


def test_external_potential_vs_oracle(gpu_ctx, oracle):
    Z, xyz = workloads.water_cluster(64)
    rng = np.random.default_rng(11)
    m = 500
    pcZ = rng.choice([1, 6, 7, 8, 11, 17], size=m)
    pc_xyz = rng.uniform(-8, 20, size=(m, 3))
    pc_q = rng.uniform(-0.8, 0.8, size=m)
    field = [0.05, -0.02, 0.1]
    P = oracle.default_parameters()
    chi, hard, rad, nq = oracle.gather(Z, P)
    _, _, prad, pnq = oracle.gather(pcZ, P)
    for basis in (1, 0):
        ref = oracle.external_potential(xyz, rad, nq, pc_xyz, pc_q, prad, pnq, field, oracle.Options(basis=basis))
        ext = _ffi.External()
        ext.num_point_charges = m
        pxyz = np.ascontiguousarray(pc_xyz); pq = np.ascontiguousarray(pc_q)
        ext.xyz, ext.charge, ext.radius, ext.n = _ptr(pxyz), _ptr(pq), _ptr(prad), _ptr(pnq)
        ext.field[0], ext.field[1], ext.field[2] = field
        v = np.empty(len(Z))
        x = np.ascontiguousarray(xyz)
        rc = gpu_ctx._lib.cheq_external_potential(gpu_ctx.handle, len(Z), _ptr(x), _ptr(rad), _ptr(nq), C.byref(ext),
                                                  0.5, basis, _ptr(v))
        assert rc == 0, gpu_ctx.last_error()
        assert np.max(np.abs(v - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))


# ---- full solves: Rappe-Goddard fixtures -----------------------------------------------------------------------
@pytest.mark.parametrize("basis", [BasisType.Sto, BasisType.Gto])
def test_rappe_goddard_parity_and_published_rows(gpu_ctx, oracle, rg_fixtures, basis):
    key = "sto" if basis is BasisType.Sto else "gto"
    s = solver(gpu_ctx, basis_type=basis)
    checked = 0
    worst = (0.0, 0.0)
    for g in rg_fixtures["groups"]:
        errs, perr = [], []
        for c in g["cases"]:
            res = s.solve(atoms_of(c["Z"], c["xyz"]), 0.0)
            ref = oracle.solve(c["Z"], np.array(c["xyz"]), 0.0, oracle.Options(basis=basis.value))
            dq, dmu = assert_parity(res, ref)
            worst = (max(worst[0], dq), max(worst[1], dmu))
            for p in c["published"]:
                val = res.charges[p["index"]]
                assert abs(round(val, 4) - p[key]) < 1e-9 or abs(val - p[key]) < 5.001e-5, (c["name"], p, val)
                perr.append(abs(round(val, 4) - p["paper"]))
                checked += 1
            errs += [abs(res.charges[i] - q) for i, q in c["expected"]]
        assert f"{np.mean(perr):.4f}" == f"{g['published_mae_' + key]:.4f}"
        if basis is BasisType.Sto:       # tests/*.rs thresholds
            assert np.mean(errs) <= g["avg_limit"] and np.max(errs) <= g["max_limit"]
    assert checked == 67
    print("worst |dq|, |dmu|:", worst)


def test_indefinite_matrix_c2h4(gpu_ctx, oracle, rg_fixtures):
    """The reference's C2H4 fixture (coincident hydrogens) is indefinite: handled by the signed factorisation."""
    c = [c for g, c in rg_cases(rg_fixtures) if c["name"] == "C2H4"][0]
    res = solver(gpu_ctx).solve(atoms_of(c["Z"], c["xyz"]), 0.0)
    assert gpu_ctx.stats()["lu_fallback_frames"] == 0
    assert_parity(res, oracle.solve(c["Z"], np.array(c["xyz"])))


def test_pivoted_lu_fallback_on_vanishing_pivot(gpu_ctx, oracle):
    """Two hydrogens placed where J_12 equals the hydrogen hardness: the leading 2 x 2 block is singular, the
    second unpivoted pivot vanishes (< 1e-4 eV) and the engine re-solves with the pivoted LU, the reference's
    own algorithm.  hydrogen_scf is off so that the hydrogens sit in the leading block."""
    P = oracle.default_parameters()
    hard, rad = P[1][1], P[1][2]
    f = lambda d: oracle.sto_integral(d / BOHR, 1, rad / BOHR, 1, rad / BOHR, 0.5) * EV - hard
    lo, hi = 0.01, 6.0
    assert f(lo) > 0 > f(hi)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(mid) > 0:
            lo = mid
        else:
            hi = mid
    d = 0.5 * (lo + hi)
    Z = [1, 1, 17]
    xyz = np.array([[0.0, 0, 0], [d, 0, 0], [0.0, 2.4, 0]])
    o = SolverOptions(hydrogen_scf=False)
    res = QEqSolver(get_default_parameters(), o, gpu_ctx).solve(atoms_of(Z, xyz), 0.0)
    assert gpu_ctx.stats()["lu_fallback_frames"] == 1
    assert_parity(res, oracle.solve(Z, xyz, 0.0, oracle_opts(oracle, o)))


@pytest.mark.parametrize("waters,scf", [(40, True), (333, True), (333, False)])
def test_pivoted_lu_path_forced(oracle, waters, scf):
    """The pivoted-LU path (the reference's own algorithm, implementation.rs:574-576) forced on ordinary systems
    (CHEQ_FORCE_LU=1): 32-column panels with partial pivoting inside 128-column outer blocks, the update right of an
    outer block on the FP64 tensor cores.  999 atoms = 9 outer blocks."""
    import os
    Z, xyz = workloads.water_cluster(waters)
    o = SolverOptions(hydrogen_scf=scf)
    ref = oracle.solve(Z, xyz, 0.0, oracle_opts(oracle, o))
    os.environ["CHEQ_FORCE_LU"] = "1"
    try:
        ctx = cheq_b200.Context(0)
    finally:
        del os.environ["CHEQ_FORCE_LU"]
    res = QEqSolver(get_default_parameters(), o, ctx).solve(atoms_of(Z, xyz), 0.0)
    assert ctx.stats()["lu_fallback_frames"] == 1
    assert_parity(res, ref)
    ctx.close()


# ---- the reference's solver unit tests (implementation.rs:620-1013) -------------------------------------------
def test_reference_solver_unit_tests(gpu_ctx):
    s = solver(gpu_ctx)
    r = s.solve([Atom(1, [0, 0, 0]), Atom(1, [0.74, 0, 0])], 0.0)
    assert abs(r.charges[0]) < 1e-6 and abs(r.charges[0] - r.charges[1]) < 1e-10
    hf = [Atom(1, [0, 0, 0]), Atom(9, [0.917, 0, 0])]
    r = s.solve(hf, 0.0)
    assert r.charges[0] > 0 > r.charges[1] and abs(sum(r.charges)) < 1e-6
    water = [Atom(8, [0, 0, 0.1173]), Atom(1, [0, 0.7572, -0.4692]), Atom(1, [0, -0.7572, -0.4692])]
    rs = s.solve(water, 0.0)
    assert rs.charges[0] < -0.1 and rs.charges[1] > 0.05 and abs(rs.charges[1] - rs.charges[2]) < 1e-6
    rg = solver(gpu_ctx, basis_type=BasisType.Gto).solve(water, 0.0)
    assert abs(rs.charges[0] - rg.charges[0]) < 0.3
    with pytest.raises(cheq_b200.NotConverged) as e:
        solver(gpu_ctx, max_iterations=1).solve([Atom(3, [0, 0, 0]), Atom(1, [1.595, 0, 0])], 0.0)
    assert e.value.max_iterations == 1 and e.value.delta > 1e-6
    a = solver(gpu_ctx, damping=DampingStrategy.fixed(0.5)).solve(hf, 0.0)
    b = solver(gpu_ctx, damping=DampingStrategy.none()).solve(hf, 0.0)
    assert a.charges[0] > 0 and b.charges[0] > 0 and abs(a.charges[0] - b.charges[0]) < 0.01
    g = solver(gpu_ctx, basis_type=BasisType.Gto).solve(hf, 0.0)
    assert g.charges[0] > 0 > g.charges[1] and abs(sum(g.charges)) < 1e-6
    # README example: neutrality to 1e-9 (lib.rs:57)
    Z, xyz = workloads.readme_water()
    r = s.solve(atoms_of(Z, xyz), 0.0)
    assert r.charges[0] < 0 and abs(sum(r.charges)) < 1e-9


def test_reference_solve_in_field_unit_tests(gpu_ctx):
    s = solver(gpu_ctx)
    hf = [Atom(1, [0, 0, 0]), Atom(9, [0.917, 0, 0])]
    a = s.solve(hf, 0.0)
    b = s.solve_in_field(hf, 0.0, ExternalPotential.new())
    assert abs(a.charges[0] - b.charges[0]) < 1e-10 and abs(a.charges[1] - b.charges[1]) < 1e-10
    co = [Atom(6, [0, 0, 0]), Atom(8, [1.2, 0, 0])]
    vac = s.solve(co, 0.0)
    f = s.solve_in_field(co, 0.0, ExternalPotential.from_point_charges([PointCharge.new(7, [3.0, 0, 0], 0.5)]))
    assert f.charges[1] < vac.charges[1] and abs(sum(f.charges)) < 1e-6
    hh = [Atom(1, [-0.37, 0, 0]), Atom(1, [0.37, 0, 0])]
    sym = s.solve_in_field(hh, 0.0, ExternalPotential.from_point_charges(
        [PointCharge.new(8, [0, 3.0, 0], 0.5), PointCharge.new(8, [0, -3.0, 0], 0.5)]))
    assert abs(sym.charges[0] - sym.charges[1]) < 1e-6
    e = s.solve_in_field(hf, 0.0, ExternalPotential.from_uniform_field([0.5, 0, 0]))
    assert e.charges[0] < a.charges[0] and abs(sum(e.charges)) < 1e-6
    both = ExternalPotential.new().with_point_charges([PointCharge.new(7, [3.0, 0, 0], 0.3)]).with_uniform_field([0.1, 0, 0])
    assert abs(sum(s.solve_in_field(co, 0.0, both).charges)) < 1e-6
    g = solver(gpu_ctx, basis_type=BasisType.Gto).solve_in_field(
        co, 0.0, ExternalPotential.from_point_charges([PointCharge.new(7, [3.0, 0, 0], 0.5)]))
    assert abs(sum(g.charges)) < 1e-6


# ---- option coverage against the oracle -----------------------------------------------------------------------
@pytest.mark.parametrize("opts", [
    dict(), dict(basis_type=BasisType.Gto), dict(hydrogen_scf=False), dict(damping=DampingStrategy.none()),
    dict(damping=DampingStrategy.fixed(0.5)), dict(damping=DampingStrategy.auto(0.9)), dict(lambda_scale=0.4913),
    dict(tolerance=1e-10), dict(tolerance=1e-3),
])
def test_options_parity_on_clusters(gpu_ctx, oracle, opts):
    o = SolverOptions(**opts)
    s = QEqSolver(get_default_parameters(), o, gpu_ctx)
    for (Z, xyz), Q in ((workloads.water_cluster(27), 0.0), (workloads.organic_chain(150), 1.0),
                        (workloads.water_cluster(8), -2.0)):
        res = s.solve(atoms_of(Z, xyz), Q)
        ref = oracle.solve(Z, xyz, Q, oracle_opts(oracle, o))
        assert_parity(res, ref)
        # (with damping < 1 the returned mix sums to Q only in the limit -- same in the reference)
        assert abs(sum(res.charges) - float(np.sum(ref.charges))) < 1e-8


def test_water_375_and_heavy_elements(gpu_ctx, oracle):
    Z, xyz = workloads.water_cluster(125)
    res = solver(gpu_ctx).solve(atoms_of(Z, xyz), 0.0)
    assert_parity(res, oracle.solve(Z, xyz))
    for Zs, box, dmin in (([11, 17] * 40 + [19, 35] * 20 + [26, 8, 8] * 10, 18.0, 2.2),      # no hydrogen: 1 solve
                          ([11, 17] * 20 + [14, 8, 8] * 8 + [1] * 12 + [6] * 10, 16.0, 2.0)):  # 22 SCF iterations
        xyz = _random_cluster(len(Zs), 5, box, dmin)
        res = solver(gpu_ctx).solve(atoms_of(Zs, xyz), 0.0)
        assert_parity(res, oracle.solve(Zs, xyz))


def test_not_converged_matches_reference_behaviour(gpu_ctx, oracle):
    """A cluster on which the reference's SCF oscillates for ever: both paths report NotConverged."""
    Zs = [3, 9] * 10 + [55, 53] * 10 + [1, 1, 8] * 15 + [79, 92, 29, 30, 46]
    xyz = _random_cluster(len(Zs), 5, 16.0, 2.2)
    o = SolverOptions(max_iterations=60)
    with pytest.raises(oracle.OracleError) as e:
        oracle.solve(Zs, xyz, 0.0, oracle_opts(oracle, o))
    assert e.value.kind == "NotConverged"
    with pytest.raises(cheq_b200.NotConverged) as e:
        QEqSolver(get_default_parameters(), o, gpu_ctx).solve(atoms_of(Zs, xyz), 0.0)
    assert e.value.max_iterations == 60


def _random_cluster(count, seed, box, dmin):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0, box, size=(20000, 3))
    keep = [pts[0]]
    for p in pts[1:]:
        if len(keep) == count:
            break
        if min(np.linalg.norm(p - k) for k in keep) > dmin:
            keep.append(p)
    assert len(keep) == count
    return np.array(keep)


def test_b3k_field_and_point_charges(gpu_ctx, oracle):
    """BASELINE config 3: 3,000-atom water cluster + uniform field + 10,000 embedding charges."""
    Z, xyz, ext = workloads.water_box_3k()
    ep = ExternalPotential.from_point_charges(
        [PointCharge(int(z), p.tolist(), float(q)) for z, p, q in zip(ext["Z"], ext["xyz"], ext["q"])]
    ).with_uniform_field(ext["field"])
    res = solver(gpu_ctx).solve_in_field(atoms_of(Z, xyz), 0.0, ep)
    ref = oracle.solve(Z, xyz, 0.0, external=ext)
    assert_parity(res, ref)


def test_batch_frames_parity(gpu_ctx, oracle):
    Z, base = workloads.organic_chain(260)
    frames = workloads.md_frames(base, 5)
    s = solver(gpu_ctx)
    q, mu, its, status = s.solve_batch(Z, frames, 0.0)
    assert np.all(status == 0)
    for f in range(len(frames)):
        ref = oracle.solve(Z, frames[f])
        assert its[f] == ref.iterations
        assert np.max(np.abs(q[f] - ref.charges)) <= Q_TOL and abs(mu[f] - ref.equilibrated_potential) <= MU_TOL
    single = s.solve(atoms_of(Z, frames[2]), 0.0)
    assert np.max(np.abs(np.array(single.charges) - q[2])) < 1e-12


@pytest.mark.parametrize("tpp", [1, 3, 4])
def test_blocked_lookahead_schedule_single_gpu(gpu_ctx, oracle, tpp):
    """The right-looking panel schedule with look-ahead (the multi-GPU factorisation, run on one GPU without
    broadcasts) gives the same answer as the oracle and as the recursive schedule."""
    ctx = cheq_b200.Context(0)
    assert ctx._lib.cheq_set_factor_mode(ctx.handle, 1, tpp) == 0
    Z, xyz = workloads.water_cluster(375)
    for kw in ({}, {"hydrogen_scf": False}, {"basis_type": BasisType.Gto}):
        o = SolverOptions(**kw)
        res = QEqSolver(get_default_parameters(), o, ctx).solve(atoms_of(Z, xyz), 0.0)
        assert_parity(res, oracle.solve(Z, xyz, 0.0, oracle_opts(oracle, o)))
        rec = QEqSolver(get_default_parameters(), o, gpu_ctx).solve(atoms_of(Z, xyz), 0.0)
        assert np.max(np.abs(np.array(res.charges) - np.array(rec.charges))) < 1e-10
    c = [c for g in __import__("json").loads((__import__("pathlib").Path(__file__).parent / "golden" /
                                              "rappe_goddard.json").read_text())["groups"] for c in g["cases"]
         if c["name"] == "C2H4"][0]
    res = solver(ctx).solve(atoms_of(c["Z"], c["xyz"]), 0.0)
    assert_parity(res, oracle.solve(c["Z"], np.array(c["xyz"])))
    ctx.close()


# ---- full size: properties that need no oracle -----------------------------------------------------------------
def test_s100k_single_factorisation_properties(gpu_ctx, oracle):
    """BASELINE config 5 size on ONE B200 (81 GB trapezoid): 100,002 atoms, one signed factorisation
    (hydrogen_scf off keeps the test to ~15 s): neutrality and stationarity on sampled rows."""
    Z, xyz = workloads.water_cluster(33334)
    n = len(Z)
    chi, hard, rad, nq = workloads.gather_parameters(Z, get_default_parameters())
    s1 = QEqSolver(get_default_parameters(), SolverOptions(hydrogen_scf=False), gpu_ctx)
    r = s1.solve_arrays(np.ascontiguousarray(xyz), chi, hard, rad, nq, 0.0)
    q = np.array(r.charges)
    assert r.iterations == 1 and abs(q.sum()) < 1e-7
    rng = np.random.default_rng(2)
    for i in rng.choice(n, size=12, replace=False):
        d = np.linalg.norm(xyz - xyz[i], axis=1) / BOHR
        row = np.array([oracle.sto_integral(d[j], int(nq[i]), rad[i] / BOHR, int(nq[j]), rad[j] / BOHR, 0.5) * EV
                        if j != i else hard[i] for j in range(n)])
        assert abs(chi[i] + row @ q + r.equilibrated_potential) < 1e-7


def test_s20k_properties(gpu_ctx, oracle):
    """BASELINE headline size (20,001 atoms): neutrality, stationarity on sampled rows, permutation equivariance."""
    Z, xyz = workloads.water_20k()
    n = len(Z)
    chi, hard, rad, nq = workloads.gather_parameters(Z, get_default_parameters())
    s1 = QEqSolver(get_default_parameters(), SolverOptions(hydrogen_scf=False), gpu_ctx)
    r = s1.solve_arrays(np.ascontiguousarray(xyz), chi, hard, rad, nq, 0.0)
    q = np.array(r.charges)
    assert r.iterations == 1 and abs(q.sum()) < 1e-8
    # stationarity: chi_i + J_ii q_i + sum_j J_ij q_j + mu = 0 on 48 sampled rows, J from the CPU oracle
    rng = np.random.default_rng(1)
    for i in rng.choice(n, size=48, replace=False):
        d = np.linalg.norm(xyz - xyz[i], axis=1) / BOHR
        row = np.array([oracle.sto_integral(d[j], int(nq[i]), rad[i] / BOHR, int(nq[j]), rad[j] / BOHR, 0.5) * EV
                        if j != i else hard[i] for j in range(n)])
        assert abs(chi[i] + row @ q + r.equilibrated_potential) < 1e-8
    # full SCF solve; permuting the atoms permutes the charges
    s = solver(gpu_ctx)
    full = s.solve_arrays(np.ascontiguousarray(xyz), chi, hard, rad, nq, 0.0)
    assert 5 <= full.iterations <= 20 and abs(sum(full.charges)) < 1e-8
    perm = rng.permutation(n)
    pr = s.solve_arrays(np.ascontiguousarray(xyz[perm]), chi[perm], hard[perm], rad[perm], nq[perm], 0.0)
    assert pr.iterations == full.iterations
    assert np.max(np.abs(np.array(pr.charges) - np.array(full.charges)[perm])) < 1e-8
    assert abs(pr.equilibrated_potential - full.equilibrated_potential) < 1e-8
