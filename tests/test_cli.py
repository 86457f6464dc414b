"""CLI mirror of /root/reference/src/bin/modules/{cli,app,io}.rs: XYZ reader, the four writers, flag defaults
(io.rs:558-773 and app.rs:77-158 are the reference's tests for these).  CPU tests use a stand-in solver; the GPU
test runs `cheq water.xyz` end to end (BASELINE config 1)."""
import io
import json
import math

import pytest

from cheq_b200 import Atom, CalculationResult, cli

WATER = "3\nWater molecule\nO 0.000000 0.000000 0.117300\nH 0.000000 0.757200 -0.469200\nH 0.000000 -0.757200 -0.469200\n"


class FakeSolver:
    def __init__(self, params, opts):
        self.opts = opts

    def solve_in_field(self, atoms, q, ext):
        n = len(atoms)
        return CalculationResult([(-0.5 if a.atomic_number == 8 else 0.25) for a in atoms], -6.5, 9)


def test_read_atoms_and_errors(tmp_path):
    f = tmp_path / "w.xyz"
    f.write_text(WATER)
    atoms, comment = cli.read_atoms(str(f))
    assert [a.atomic_number for a in atoms] == [8, 1, 1] and comment == "Water molecule"
    assert atoms[1].position == [0.0, 0.7572, -0.4692]
    (tmp_path / "bad.xyz").write_text("3\nc\nO 0 0 0\nH 0 0 1\n")
    with pytest.raises(cli.CliError) as e:
        cli.read_atoms(str(tmp_path / "bad.xyz"))
    assert "Expected 3 atoms, got 2" in str(e.value)
    (tmp_path / "el.xyz").write_text("1\nc\nXx 0 0 0\n")
    with pytest.raises(cli.CliError) as e:
        cli.read_atoms(str(tmp_path / "el.xyz"))
    assert "Unknown element: Xx" in str(e.value)
    with pytest.raises(cli.CliError):
        cli.read_atoms(str(tmp_path / "missing.xyz"))
    # element tokens: numbers and case-insensitive symbols (io.rs:90-214)
    assert cli.parse_element("8") == 8 and cli.parse_element("cl") == 17 and cli.parse_element("CL") == 17
    assert cli.parse_element("h") == 1 and cli.parse_element("Zz") is None
    frames = tmp_path / "traj.xyz"
    frames.write_text(WATER + WATER.replace("0.117300", "0.120000"))
    fr = cli.read_frames(str(frames))
    assert len(fr) == 2 and fr[1][0][0].position[2] == 0.12


def test_writers():
    atoms = [Atom(8, [0.0, 0.0, 0.1173]), Atom(1, [0.0, 0.7572, -0.4692])]
    res = CalculationResult([-0.5, 0.5], -6.54321, 9)
    w = io.StringIO()
    cli.write_csv(w, atoms, res, 4)
    assert w.getvalue().splitlines() == ["index,element,x,y,z,charge", "0,O,0.0000,0.0000,0.1173,-0.5000",
                                         "1,H,0.0000,0.7572,-0.4692,0.5000"]
    w = io.StringIO()
    cli.write_xyz(w, atoms, res, " comment ", 3)
    lines = w.getvalue().splitlines()
    assert lines[0] == "2" and lines[1] == "comment | QEq charges | iterations: 9 | potential: -6.543"
    assert lines[2] == "O 0.000 0.000 0.117 -0.500"
    w = io.StringIO()
    cli.write_json(w, atoms, res, 6)
    doc = json.loads(w.getvalue())
    assert doc["iterations"] == 9 and doc["atoms"][1]["element"] == "H" and doc["total_charge"] == 0.0
    assert doc["equilibrated_potential"] == -6.54321 and doc["atoms"][0]["position"] == [0.0, 0.0, 0.1173]
    w = io.StringIO()
    cli.write_pretty(w, atoms, res, 6, "water.xyz")
    text = w.getvalue()
    assert "Cheq Charge Equilibration Results" in text and "Converged Iterations:" in text and "Charge (e)" in text
    assert "-6.543210 eV" in text and "water.xyz" in text


def test_flag_defaults_and_run_with_stand_in_solver(tmp_path, capsys):
    a = cli.build_parser().parse_args(["x.xyz"])
    o = cli.options_from_args(a)
    assert o.max_iterations == 100 and o.tolerance == 1e-6 and o.lambda_scale == 0.5 and o.hydrogen_scf is True
    assert o.basis_type.name == "Sto" and o.damping.kind == "auto" and o.damping.value == 0.4
    assert a.format == "pretty" and a.precision == 6 and a.total_charge == 0.0
    a = cli.build_parser().parse_args(["x.xyz", "--basis", "gto", "--damping", "fixed", "--damping-factor", "0.5",
                                       "--hydrogen-scf", "false", "-q", "1", "-f", "json", "-p", "3"])
    o = cli.options_from_args(a)
    assert o.basis_type.name == "Gto" and o.damping.kind == "fixed" and o.damping.value == 0.5 and not o.hydrogen_scf
    f = tmp_path / "w.xyz"
    f.write_text(WATER)
    out = tmp_path / "out.csv"
    rc = cli.run([str(f), "-f", "csv", "-o", str(out)], solver_factory=FakeSolver)
    assert rc == 0 and out.read_text().splitlines()[1].startswith("0,O,0.000000,0.000000,0.117300,-0.500000")
    rc = cli.run([str(tmp_path / "nope.xyz")], solver_factory=FakeSolver)
    assert rc == 1 and "Error:" in capsys.readouterr().err


@pytest.mark.gpu
def test_cheq_water_xyz_end_to_end(tmp_path, gpu_ctx, oracle):
    """BASELINE config 1: `cheq water.xyz`, default STO parameters, total charge 0."""
    import numpy as np
    from cheq_b200 import workloads
    Z, xyz = workloads.readme_water()
    f = tmp_path / "water.xyz"
    f.write_text("3\nREADME water\n" + "".join(f"{'O' if z == 8 else 'H'} {p[0]:.10f} {p[1]:.10f} {p[2]:.10f}\n"
                                              for z, p in zip(Z, xyz)))
    out = tmp_path / "water.json"
    rc = cli.run([str(f), "-f", "json", "-p", "10", "-o", str(out)])
    assert rc == 0
    doc = json.loads(out.read_text())
    atoms, _ = cli.read_atoms(str(f))
    ref = oracle.solve(Z, np.array([a.position for a in atoms]), 0.0, oracle.Options(max_iterations=100))
    assert doc["iterations"] == ref.iterations
    assert max(abs(a["charge"] - q) for a, q in zip(doc["atoms"], ref.charges)) < 1e-8
    assert abs(doc["equilibrated_potential"] - ref.equilibrated_potential) < 1e-8
    # trajectory mode = per-frame solves
    traj = tmp_path / "traj.xyz"
    traj.write_text(f.read_text() * 3)
    out2 = tmp_path / "traj.csv"
    assert cli.run([str(traj), "--trajectory", "-f", "csv", "-o", str(out2)]) == 0
    assert len(out2.read_text().splitlines()) == 3 * 4


@pytest.mark.gpu
def test_cli_field_and_point_charges_end_to_end(tmp_path, gpu_ctx, oracle):
    """`--field` and `--point-charges` drive QEqSolver::solve_in_field
    (/root/reference/src/solver/implementation.rs:239-266) from the command line; result vs the CPU oracle."""
    import numpy as np
    from cheq_b200 import workloads
    Z, xyz = workloads.water_cluster(27)
    sym = {8: "O", 1: "H", 6: "C", 7: "N"}
    f = tmp_path / "cluster.xyz"
    f.write_text(f"{len(Z)}\ncluster\n" + "".join(f"{sym[z]} {p[0]:.10f} {p[1]:.10f} {p[2]:.10f}\n" for z, p in zip(Z, xyz)))
    rng = np.random.default_rng(3)
    pcZ = [8, 1, 7, 6, 1, 8]
    pc_xyz = xyz.mean(axis=0) + rng.normal(size=(len(pcZ), 3)) * 3.0 + np.array([14.0, 0.0, 0.0])
    pc_q = np.array([-0.6, 0.3, -0.2, 0.1, 0.3, 0.1])
    pcf = tmp_path / "charges.txt"
    pcf.write_text("".join(f"{sym[z]} {p[0]:.10f} {p[1]:.10f} {p[2]:.10f} {q:.10f}\n" for z, p, q in zip(pcZ, pc_xyz, pc_q)))
    field = [0.05, -0.02, 0.1]
    atoms, _ = cli.read_atoms(str(f))
    axyz = np.array([a.position for a in atoms])
    pcs = cli.read_point_charges(str(pcf))
    ext = {"Z": [pc.atomic_number for pc in pcs], "xyz": np.array([pc.position for pc in pcs]),
           "q": np.array([pc.charge for pc in pcs]), "field": field}
    for args, e in ((["--field", *map(str, field)], {"Z": [], "xyz": np.zeros((0, 3)), "q": [], "field": field}),
                    (["--point-charges", str(pcf)], dict(ext, field=[0.0, 0.0, 0.0])),
                    (["--point-charges", str(pcf), "--field", *map(str, field)], ext)):
        out = tmp_path / "out.json"
        assert cli.run([str(f), "-f", "json", "-p", "10", "-o", str(out), *args]) == 0
        doc = json.loads(out.read_text())
        ref = oracle.solve(Z, axyz, 0.0, oracle.Options(max_iterations=100), external=e)
        assert doc["iterations"] == ref.iterations
        assert max(abs(a["charge"] - q) for a, q in zip(doc["atoms"], ref.charges)) < 1e-8
        assert abs(doc["equilibrated_potential"] - ref.equilibrated_potential) < 1e-8
    vac = oracle.solve(Z, axyz, 0.0, oracle.Options(max_iterations=100))
    assert max(abs(a["charge"] - q) for a, q in zip(doc["atoms"], vac.charges)) > 1e-4   # the field really acts
