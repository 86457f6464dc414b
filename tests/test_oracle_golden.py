"""The oracle against every golden vector the reference holds for the hot path (CPU only).

  * BENCHMARKS.md:20-89 -- 67 rows x {STO, GTO} of cheq's own outputs at 4 decimals, and the 4 category MAEs
  * gto.rs:182-197 -- two GTO known answers at 1e-8 eV
  * tests/{alkali,clusters,hydrogen,polymers}.rs -- group thresholds (STO, total charge 0)
  * sto.rs:57-109, gto.rs:154-228 -- behavioural invariants of the pair integrals
  * tests/golden/sto_truth.json -- 40-digit values of the integral's definition (oracle/sto_truth_mp.py)
"""
import mpmath as mp
import numpy as np
import pytest

from conftest import rg_cases

BOHR = 0.529177210903
EV = 27.211386245988


def test_gto_known_answers(oracle):
    lam = 0.4913
    r_h = 0.371 / BOHR
    v = oracle.gto_integral(0.74 / BOHR, 1, r_h, 1, r_h, lam) * EV
    assert abs(v - 14.02504752) < 1e-8
    v = oracle.gto_integral(1.128 / BOHR, 2, 0.759 / BOHR, 2, 0.669 / BOHR, lam) * EV
    assert abs(v - 5.837573337) < 1e-8


def test_sto_matches_definition_to_fp64(oracle, sto_truth):
    worst = 0.0
    for r in sto_truth["rows"]:
        z1, z2, R = float.fromhex(r["zeta1"]), float.fromhex(r["zeta2"]), float.fromhex(r["R_bohr"])
        J = oracle.sto_coulomb(R, r["n1"], z1, r["n2"], z2)
        t = mp.mpf(r["J_hartree"])
        worst = max(worst, float(abs((mp.mpf(J) - t) / t)))
    assert worst < 5e-16, worst


def test_sto_invariants(oracle):
    lam = 0.5
    r_h, r_c = 0.371 / BOHR, 0.759 / BOHR
    a = oracle.sto_integral(2.0, 1, r_h, 2, r_c, lam)
    b = oracle.sto_integral(2.0, 2, r_c, 1, r_h, lam)
    assert abs(a - b) < 1e-12                                   # sto.rs:57-69
    assert abs(oracle.sto_integral(100.0, 1, r_h, 1, r_h, lam) - 0.01) < 1e-6   # sto.rs:72-81
    j0 = oracle.sto_integral(0.0, 1, r_h, 1, r_h, lam)
    j1 = oracle.sto_integral(1e-10, 1, r_h, 1, r_h, lam)
    assert np.isfinite(j0) and abs(j0 - j1) < 1e-6               # sto.rs:84-94
    g = oracle.gto_integral(2.0, 1, r_h, 2, r_c, lam)
    assert abs(a - g) < 0.25                                     # sto.rs:97-109
    zeta = lam * 3 / (2 * r_h)
    assert abs(j0 - 5 * zeta / 8) < 1e-14                        # 1s-1s closed form at R = 0


def test_gto_invariants(oracle):
    lam = 0.5
    r_h, r_c = 0.371 / BOHR, 0.759 / BOHR
    assert oracle.gto_integral(2.0, 1, r_h, 2, r_c, lam) == oracle.gto_integral(2.0, 2, r_c, 1, r_h, lam)
    assert abs(oracle.gto_integral(1000.0, 1, r_h, 2, r_c, lam) - 1e-3) < 1e-9     # gto.rs:169-179
    lim = oracle.gto_integral(0.0, 1, r_h, 1, r_h, lam)
    near = oracle.gto_integral(1e-8, 1, r_h, 1, r_h, lam)
    assert abs(lim - near) < 1e-6                                                  # gto.rs:154-166
    assert oracle.gto_integral(2.0, 1, r_h, 1, r_h, 0.3) < oracle.gto_integral(2.0, 1, r_h, 1, r_h, 0.6)


@pytest.mark.parametrize("basis,key", [(1, "sto"), (0, "gto")])
def test_benchmarks_md_rows(oracle, rg_fixtures, basis, key):
    """All 67 published rows per basis to the 4 printed decimals, and the category MAEs."""
    checked = 0
    for g in rg_fixtures["groups"]:
        errs = []
        for c in g["cases"]:
            res = oracle.solve(c["Z"], np.array(c["xyz"]), 0.0, oracle.Options(basis=basis))
            for p in c["published"]:
                val = res.charges[p["index"]]
                assert abs(round(val, 4) - p[key]) < 1e-9 or abs(val - p[key]) < 5.001e-5, (c["name"], p, val)
                errs.append(abs(round(val, 4) - p["paper"]))
                checked += 1
        assert f"{np.mean(errs):.4f}" == f"{g['published_mae_' + key]:.4f}", g["name"]
    assert checked == 67


def test_group_thresholds(oracle, rg_fixtures):
    """tests/common/mod.rs:9-93 with the limits of each group file (STO, Q = 0)."""
    for g in rg_fixtures["groups"]:
        errs = []
        for c in g["cases"]:
            res = oracle.solve(c["Z"], np.array(c["xyz"]), 0.0)
            errs += [abs(res.charges[i] - q) for i, q in c["expected"]]
        assert np.mean(errs) <= g["avg_limit"], g["name"]
        assert np.max(errs) <= g["max_limit"], g["name"]


def test_solver_behaviour(oracle):
    """implementation.rs:620-787 behavioural tests on the oracle."""
    r = oracle.solve([1, 1], np.array([[0, 0, 0], [0.74, 0, 0]], float))
    assert abs(r.charges[0]) < 1e-6 and abs(r.charges[0] - r.charges[1]) < 1e-10
    hf = np.array([[0, 0, 0], [0.917, 0, 0]], float)
    r = oracle.solve([1, 9], hf)
    assert r.charges[0] > 0 > r.charges[1] and abs(sum(r.charges)) < 1e-6
    with pytest.raises(oracle.OracleError) as e:
        oracle.solve([3, 1], np.array([[0, 0, 0], [1.595, 0, 0]], float), opt=oracle.Options(max_iterations=1))
    assert e.value.kind == "NotConverged"
    with pytest.raises(oracle.OracleError) as e:
        oracle.solve([], np.zeros((0, 3)))
    assert e.value.kind == "NoAtoms"
    with pytest.raises(oracle.OracleError) as e:
        oracle.solve([118], np.zeros((1, 3)))
    assert e.value.kind == "ParameterNotFound" and e.value.info["z"] == 118
    a = oracle.solve([1, 9], hf, opt=oracle.Options(damping=("fixed", 0.5)))
    b = oracle.solve([1, 9], hf, opt=oracle.Options(damping=("none",)))
    assert abs(a.charges[0] - b.charges[0]) < 0.01
    # SURVEY.md App. B pinned probe values
    import math
    h = math.radians(104.5) / 2
    w = np.array([[0, 0, 0], [0.958 * math.cos(h), 0.958 * math.sin(h), 0], [0.958 * math.cos(h), -0.958 * math.sin(h), 0]])
    r = oracle.solve([8, 1, 1], w)
    assert r.iterations == 9 and abs(r.equilibrated_potential + 6.721479) < 1e-6 and abs(r.charges[1] - 0.3251) < 5e-5


def test_solve_in_field_behaviour(oracle):
    """implementation.rs:817-1013."""
    co = np.array([[0, 0, 0], [1.2, 0, 0]], float)
    vac = oracle.solve([6, 8], co)
    same = oracle.solve([6, 8], co, external={"Z": [], "xyz": np.zeros((0, 3)), "q": [], "field": [0, 0, 0]})
    assert np.max(np.abs(vac.charges - same.charges)) < 1e-10
    f = oracle.solve([6, 8], co, external={"Z": [7], "xyz": np.array([[3.0, 0, 0]]), "q": [0.5], "field": [0, 0, 0]})
    assert f.charges[1] < vac.charges[1] and abs(sum(f.charges)) < 1e-6
    hh = np.array([[-0.37, 0, 0], [0.37, 0, 0]], float)
    s = oracle.solve([1, 1], hh, external={"Z": [8, 8], "xyz": np.array([[0, 3.0, 0], [0, -3.0, 0]]), "q": [0.5, 0.5],
                                           "field": [0, 0, 0]})
    assert abs(s.charges[0] - s.charges[1]) < 1e-6
    hf = np.array([[0, 0, 0], [0.917, 0, 0]], float)
    v = oracle.solve([1, 9], hf)
    e = oracle.solve([1, 9], hf, external={"Z": [], "xyz": np.zeros((0, 3)), "q": [], "field": [0.5, 0, 0]})
    assert e.charges[0] < v.charges[0]
    with pytest.raises(oracle.OracleError) as err:
        oracle.solve([6], np.zeros((1, 3)), external={"Z": [200], "xyz": np.array([[3.0, 0, 0]]), "q": [0.5],
                                                      "field": [0, 0, 0]})
    assert err.value.info["z"] == 200
