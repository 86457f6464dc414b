"""CPU tests of the host side: API mirror, TOML loader, C-ABI library surface, STO table construction,
and the algebraic reformulation the CUDA path uses.  No compute call reaches a GPU here."""
import ctypes as C
import re
from pathlib import Path

import mpmath as mp
import numpy as np
import pytest

import cheq_b200
from cheq_b200 import (Atom, BasisType, DampingStrategy, ExternalPotential, Parameters, PointCharge, QEqSolver,
                       SolverOptions, _ffi, get_default_parameters)
from conftest import ROOT, rg_cases


# ---- params.rs:350-501 ---------------------------------------------------------------------------------------
TOML = """
[elements]
"1" = { chi = 2.20, j = 13.60, radius = 0.37, n = 1 }
"Fe" = { chi = 1.83, j = 7.90, radius = 1.26, n = 4 }
"O" = { chi = 3.44, j = 13.62, radius = 0.60, n = 2 }
"""


def test_load_from_str_mixed_keys():
    p = Parameters.load_from_str(TOML)
    assert set(p.elements) == {1, 26, 8}
    assert p.elements[26].hardness == 7.90 and p.elements[26].principal_quantum_number == 4
    assert p.elements[1].electronegativity == 2.20 and p.elements[8].radius == 0.60


def test_load_errors(tmp_path):
    with pytest.raises(cheq_b200.DeserializationError):
        Parameters.load_from_str("this is not valid toml")
    with pytest.raises(cheq_b200.DeserializationError) as e:
        Parameters.load_from_str('[elements]\n"InvalidKey" = { chi = 1.0, j = 1.0, radius = 1.0, n = 1 }')
    assert "invalid element key: 'InvalidKey'" in str(e.value)
    with pytest.raises(cheq_b200.DeserializationError):
        Parameters.load_from_str('[elements]\n"1" = { chi = 2.20, j = 13.60, radius = 0.37 }')
    with pytest.raises(cheq_b200.IoError):
        Parameters.load_from_file(tmp_path / "non_existent_file.toml")
    f = tmp_path / "p.toml"
    f.write_text(TOML)
    assert Parameters.load_from_file(f).elements == Parameters.load_from_str(TOML).elements
    assert Parameters.new().elements == {} and Parameters().elements == {}


def test_default_parameters_cached_and_complete():
    a, b = get_default_parameters(), get_default_parameters()
    assert a is b                                   # lib.rs:135-157
    assert len(a.elements) == 103
    assert a.elements[1].hardness == 13.8904 and a.elements[8].radius == 0.669
    assert a.elements[55].principal_quantum_number == 6


# ---- types.rs:294-416 -----------------------------------------------------------------------------------------
def test_external_potential_builders():
    assert ExternalPotential.new().is_empty() and ExternalPotential().is_empty()
    e = ExternalPotential.from_point_charges([PointCharge.new(8, [1.0, 0, 0], -0.5), PointCharge(1, [2.0, 0, 0], 0.25)])
    assert not e.is_empty() and e.num_point_charges() == 2 and e.point_charges()[1].charge == 0.25
    f = ExternalPotential.from_uniform_field([0.1, 0.2, 0.3])
    assert not f.is_empty() and f.uniform_field() == [0.1, 0.2, 0.3] and f.num_point_charges() == 0
    g = ExternalPotential.new().with_point_charges([PointCharge(6, [0, 0, 0], 0.1)]).with_uniform_field([0, 0, 1.0])
    assert g.num_point_charges() == 1 and g.uniform_field()[2] == 1.0
    assert not ExternalPotential.from_uniform_field([0, 0, 0.001]).is_empty()
    assert not ExternalPotential.from_point_charges([PointCharge(1, [0, 0, 0], 0.0)]).is_empty()


def test_options_defaults():
    o = SolverOptions.default()
    assert (o.tolerance, o.max_iterations, o.lambda_scale, o.hydrogen_scf) == (1e-6, 2000, 0.5, True)
    assert o.basis_type is BasisType.Sto and o.damping == DampingStrategy.auto(0.4)
    m = cheq_b200.solver.make_options(SolverOptions(basis_type=BasisType.Gto, damping=DampingStrategy.fixed(0.5)))
    assert (m.basis, m.damping_kind, m.damping_value) == (0, 1, 0.5)


def test_host_side_errors_need_no_gpu():
    """NoAtoms and ParameterNotFound are decided before the FFI call (implementation.rs:180-182, 348-362)."""
    s = QEqSolver.new(get_default_parameters())
    with pytest.raises(cheq_b200.NoAtoms):
        s.solve([], 0.0)
    with pytest.raises(cheq_b200.ParameterNotFound) as e:
        s.solve([Atom(118, [0, 0, 0])], 0.0)
    assert e.value.atomic_number == 118 and "118" in str(e.value)
    ext = ExternalPotential.from_point_charges([PointCharge(200, [3.0, 0, 0], 0.5)])
    with pytest.raises(cheq_b200.ParameterNotFound) as e:
        s.solve_in_field([Atom(6, [0, 0, 0])], 0.0, ext)
    assert e.value.atomic_number == 200
    msg = str(cheq_b200.NotConverged(100, 1.234e-3))
    assert msg == "SCF failed to converge after 100 iterations. Final charge difference: 1.23e-3" or "1.23e-03" in msg


# ---- C ABI surface -------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "cheq_b200.h").read_text()
    declared = set(re.findall(r"\b(cheq_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    lib = _ffi.lib()
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.Options) == 32
    assert C.sizeof(_ffi.External) == 8 + 4 * 8 + 24
    assert C.sizeof(_ffi.Stats) == 7 * 8 + 4 * 8 + 2 * 4 + 2 * 8 + 8 + 5 * 8 + 4 * 8 + 2 * 8 + 4 * 8   # + the four krylov_*, two matfree_* and four tc_gemm_* fields
    o = _ffi.Options()
    _ffi.lib().cheq_options_default(C.byref(o))
    assert (o.tolerance, o.lambda_scale, o.damping_value, o.max_iterations) == (1e-6, 0.5, 0.4, 2000)
    assert (o.hydrogen_scf, o.basis, o.damping_kind) == (1, 1, 2)


# ---- STO pair tables (host construction, fp64 evaluation as on the device) -----------------------------------
def _table_eval(n1, z1, n2, z2, R):
    kind, terms, coefs, err = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
    v = _ffi.lib().cheq_host_sto_table_eval(n1, z1, n2, z2, R, C.byref(kind), C.byref(terms), C.byref(coefs),
                                            C.byref(err))
    return v, kind.value, terms.value, coefs.value, err.value


def test_sto_tables_match_definition(sto_truth):
    worst = 0.0
    for r in sto_truth["rows"]:
        z1, z2, R = float.fromhex(r["zeta1"]), float.fromhex(r["zeta2"]), float.fromhex(r["R_bohr"])
        v, kind, terms, coefs, err = _table_eval(r["n1"], z1, r["n2"], z2, R)
        assert err <= 1e-15, (r, kind, err)
        t = mp.mpf(r["J_hartree"])
        rel = float(abs((mp.mpf(v) - t) / t))
        # (1 - F) / R loses digits only for R << 1 bohr where 1 - F is tiny; physical distances are > 1 bohr
        assert rel < (5e-14 if R < 0.5 else 5e-15), (r, rel)
        worst = max(worst, rel)
    assert worst < 5e-14


def test_sto_tables_vs_oracle_on_organic_and_degenerate_pairs(oracle):
    P = oracle.default_parameters()
    lam = 0.5

    def zeta(Z):
        _, _, r, n = P[Z]
        return n, lam * (2.0 * n + 1.0) / (2.0 * (r / 0.529177210903))
    pairs = [(1, 1), (1, 6), (1, 7), (1, 8), (6, 6), (6, 7), (6, 8), (7, 7), (7, 8), (8, 8), (24, 25), (5, 23),
             (27, 42), (11, 17), (55, 53), (26, 29), (70, 71)]
    for a, b in pairs:
        n1, z1 = zeta(a)
        n2, z2 = zeta(b)
        for R in np.geomspace(0.6, 60.0, 40):
            v = _table_eval(n1, z1, n2, z2, float(R))[0]
            ref = oracle.sto_coulomb(float(R), n1, z1, n2, z2)
            assert abs(v - ref) <= 4e-15 * abs(ref) + 1e-300, (a, b, R, v, ref)
        # symmetry (sto.rs:57-69)
        assert abs(_table_eval(n1, z1, n2, z2, 2.0)[0] - _table_eval(n2, z2, n1, z1, 2.0)[0]) < 1e-12


def test_sto_table_invalid_inputs_are_nan():
    assert np.isnan(_table_eval(0, 1.0, 1, 1.0, 2.0)[0])
    assert np.isnan(_table_eval(1, 0.0, 1, 1.0, 2.0)[0])


# ---- the reformulated algebra reproduces the bordered LU ----------------------------------------------------
def test_engine_algebra_model_matches_oracle(oracle, rg_fixtures):
    from engine_model import solve_model
    P = oracle.default_parameters()
    for g, c in rg_cases(rg_fixtures):
        xyz = np.array(c["xyz"])
        chi, hard, rad, nq = oracle.gather(c["Z"], P)
        ref = oracle.solve(c["Z"], xyz)
        M, rhs, hidx = oracle.build_system(xyz, chi, hard, rad, nq, 0.0, None, oracle.Options())
        n = len(chi)
        q, mu, it = solve_model(M[:n, :n].copy(), chi, hard, nq, 0.0)
        assert it == ref.iterations
        assert np.max(np.abs(q - ref.charges)) < 1e-12 and abs(mu - ref.equilibrated_potential) < 1e-12


def test_stale_factor_model_matches_oracle(oracle, rg_fixtures):
    """The stale-factor SCF iterations (frozen hydrogen factor + preconditioned GMRES on the reduced bordered system,
    kernels_krylov.cu) follow the reference's trajectory: same iteration count, charges to 1e-11."""
    from cheq_b200 import workloads as W
    from engine_model import solve_model_stale
    P = oracle.default_parameters()
    cases = [(c["Z"], np.array(c["xyz"])) for g, c in rg_cases(rg_fixtures)
             if sum(1 for z in c["Z"] if z in (1, 2)) >= 4][:6]
    Zw, xw = W.water_cluster(60)
    cases.append((Zw, xw))
    Zo, xo = W.organic_chain(150)
    cases.append((Zo, xo))
    used = 0
    for Z, xyz in cases:
        chi, hard, rad, nq = oracle.gather(Z, P)
        ref = oracle.solve(Z, xyz)
        M, rhs, hidx = oracle.build_system(xyz, chi, hard, rad, nq, 0.0, None, oracle.Options())
        n = len(chi)
        trace = []
        q, mu, it = solve_model_stale(M[:n, :n].copy(), chi, hard, nq, 0.0, trace=trace)
        assert it == ref.iterations
        assert np.max(np.abs(q - ref.charges)) < 1e-11 and abs(mu - ref.equilibrated_potential) < 1e-11
        used += sum(m for m, _ in trace)
        assert all(steps <= 40 for m, steps in trace if m)
    assert used > 10   # the frozen-factor branch really ran


def test_c2h4_fixture_is_indefinite(oracle, rg_fixtures):
    """Documents why the factorisation carries pivot signs: the reference's own C2H4 geometry is not SPD."""
    P = oracle.default_parameters()
    c = [c for g, c in rg_cases(rg_fixtures) if c["name"] == "C2H4"][0]
    chi, hard, rad, nq = oracle.gather(c["Z"], P)
    M, _, _ = oracle.build_system(np.array(c["xyz"]), chi, hard, rad, nq, 0.0, None, oracle.Options())
    assert np.linalg.eigvalsh(M[:6, :6])[0] < 0


def test_workload_generators_are_seeded_and_physical():
    from cheq_b200 import workloads as W
    Z1, x1 = W.water_cluster(27)
    Z2, x2 = W.water_cluster(27)
    assert Z1 == Z2 and np.array_equal(x1, x2) and len(Z1) == 81
    d = np.linalg.norm(x1[:, None] - x1[None], axis=-1) + np.eye(81) * 10
    assert d.min() > 0.9
    Z, xyz, ext = W.water_box_3k()
    assert len(Z) == 3000 and len(ext["q"]) == 10000 and abs(ext["q"].sum()) < 1e-9
    Zo, xo = W.organic_chain(200)
    assert len(Zo) == 200 and set(Zo) <= {1, 6, 7, 8}
    fr = W.md_frames(xo, 3)
    assert fr.shape == (3, 200, 3) and not np.array_equal(fr[0], fr[1])


# ---- multi-GPU entry points: argument checking needs no GPU ----------------------------------------------------
def test_dist_entry_points_reject_bad_arguments_without_gpu():
    lib = _ffi.lib()
    assert lib.cheq_ctx_dist_init(None, 0, 2, None) == _ffi.INVALID_ARGUMENT
    assert lib.cheq_ctx_dist_info(None, None, None) == _ffi.INVALID_ARGUMENT
    assert lib.cheq_set_factor_mode(None, 0, 4) == _ffi.INVALID_ARGUMENT
    assert lib.cheq_dist_unique_id(None) == _ffi.INVALID_ARGUMENT
    assert lib.cheq_dbg_potrf_cycles(None, None, 0) == _ffi.INVALID_ARGUMENT
    owner = np.zeros(5, dtype=np.int32)
    assert lib.cheq_dist_plan(2, 3, 2, 3, owner.ctypes.data_as(C.c_void_p), None) == _ffi.OK
    assert owner.tolist() == [0, 0, 1, 1, 2]       # X block: one 2-tile panel; H block: panels of 2 + 1 tiles


def test_join_without_process_group_is_a_no_op():
    """cheq_b200.distributed.join leaves a context alone when torch.distributed is not initialised (single GPU)."""
    from cheq_b200 import distributed

    class FakeCtx:
        pass
    distributed.join(FakeCtx())                    # must not touch the context


def test_s100k_generator_is_seeded_and_physical():
    from scipy.spatial import cKDTree

    from cheq_b200 import workloads as W
    Z, xyz = W.solvated_100k(n_solute=400, n_waters=1200)
    Z2, xyz2 = W.solvated_100k(n_solute=400, n_waters=1200)
    assert Z == Z2 and np.array_equal(xyz, xyz2) and len(Z) == 400 + 3 * 1200
    assert Z[400:403] == [8, 1, 1] and set(Z[:400]) <= {1, 6, 7, 8}
    d, _ = cKDTree(xyz).query(xyz, k=2)
    assert d[:, 1].min() > 0.9                      # minimum inter-atomic distance


# ---- schedule of the pivoted-LU fallback (kernels_lu.cu) -----------------------------------------------------------
def test_two_level_lu_schedule_model():
    """tests/lu_model.py restates launch_lu_solve's kernel sequence; on a symmetric indefinite matrix that needs
    pivoting in (nearly) every column it must agree with LAPACK."""
    import lu_model
    rng = np.random.default_rng(7)
    for n in (128, 384):
        A = rng.normal(size=(n, n))
        A = A + A.T
        A[0, 0] = 0.0                      # a vanishing leading pivot, the case the fallback exists for
        b = rng.normal(size=n)
        x = lu_model.solve(A, b)
        assert np.max(np.abs(x - np.linalg.solve(A, b))) < 1e-10
