"""`python -m cheq_b200 INPUT [options]` -- the reference CLI's surface.

Mirrors /root/reference/src/bin/modules/cli.rs:25-146 (flags, defaults incl. the CLI's max_iterations = 100),
app.rs:8-75 (flow) and io.rs:8-432 (XYZ reader, pretty/xyz/csv/json writers, same field order and number
formatting).  Additions the reference library supports but its CLI does not expose: `--field EX EY EZ` and
`--point-charges FILE` (lines `element x y z charge`), and `--trajectory` for multi-frame XYZ files of one
topology (solved as one batch on the GPU).
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

from . import (Atom, BasisType, CheqError, DampingStrategy, ExternalPotential, Parameters, PointCharge, QEqSolver,
               SolverOptions, get_default_parameters)
from .params import _SYMBOLS, _SYMBOL_TO_Z


class CliError(Exception):
    pass


def parse_element(token: str):
    """io.rs:90-214: an atomic number or a case-insensitive symbol."""
    if token.isascii() and token.isdigit() and int(token) <= 255:
        return int(token)
    t = token.capitalize() if len(token) > 1 else token.upper()
    return _SYMBOL_TO_Z.get(t)


def atomic_number_to_symbol(z: int) -> str:
    return _SYMBOLS[z - 1] if 1 <= z <= len(_SYMBOLS) else "??"


def _parse_frame(lines, start, source):
    if start >= len(lines):
        raise CliError(f"Failed to parse XYZ '{source}': Missing number of atoms line")
    head = lines[start]
    try:
        n = int(head.strip())
        if n < 0:
            raise ValueError
    except ValueError:
        raise CliError(f"Failed to parse XYZ '{source}': Invalid number of atoms: {head}") from None
    if start + 1 >= len(lines):
        raise CliError(f"Failed to parse XYZ '{source}': Missing comment line")
    comment = lines[start + 1]
    atoms = []
    for i in range(n):
        if start + 2 + i >= len(lines):
            break
        parts = lines[start + 2 + i].split()
        if len(parts) < 4:
            raise CliError(f"Failed to parse XYZ '{source}': Line {i + 3}: expected at least 4 fields, got {len(parts)}")
        z = parse_element(parts[0])
        if z is None:
            raise CliError(f"Failed to parse XYZ '{source}': Unknown element: {parts[0]}")
        xyz = []
        for name, tok in zip("xyz", parts[1:4]):
            try:
                xyz.append(float(tok))
            except ValueError:
                raise CliError(f"Failed to parse XYZ '{source}': Invalid {name} coordinate: {tok}") from None
        atoms.append(Atom(z, xyz))
    if len(atoms) != n:
        raise CliError(f"Failed to parse XYZ '{source}': Expected {n} atoms, got {len(atoms)}")
    return atoms, comment, start + 2 + n


def read_atoms(input_spec: str):
    """io.rs:8-88: first frame of an XYZ file ('-' = stdin) -> (atoms, comment)."""
    frames = read_frames(input_spec, first_only=True)
    return frames[0]


def read_frames(input_spec: str, first_only: bool = False):
    try:
        text = sys.stdin.read() if input_spec == "-" else Path(input_spec).read_text()
    except OSError as e:
        raise CliError(f"I/O error at path '{input_spec}': {e}") from e
    lines = text.splitlines()
    frames, pos = [], 0
    while True:
        atoms, comment, pos = _parse_frame(lines, pos, input_spec)
        frames.append((atoms, comment))
        if first_only:
            break
        while pos < len(lines) and not lines[pos].strip():
            pos += 1
        if pos >= len(lines):
            break
    return frames


def read_point_charges(path: str):
    out = []
    for ln, line in enumerate(Path(path).read_text().splitlines(), 1):
        line = line.split("#")[0].strip()
        if not line:
            continue
        parts = line.split()
        if len(parts) < 5:
            raise CliError(f"Point-charge file '{path}' line {ln}: expected `element x y z charge`")
        z = parse_element(parts[0])
        if z is None:
            raise CliError(f"Point-charge file '{path}' line {ln}: Unknown element: {parts[0]}")
        out.append(PointCharge(z, [float(parts[1]), float(parts[2]), float(parts[3])], float(parts[4])))
    return out


# ---- writers (io.rs:229-432) ------------------------------------------------------------------------------------
def _table(rows, titles=None, align=None, inner=True):
    cols = len(rows[0]) if rows else len(titles)
    cells = ([titles] if titles else []) + rows
    width = [max(len(str(r[c])) for r in cells) for c in range(cols)]
    align = align or ["l"] * cols

    def fmt(r, center=False):
        out = []
        for c in range(cols):
            s = str(r[c])
            out.append(s.center(width[c]) if center else s.rjust(width[c]) if align[c] == "r" else s.ljust(width[c]))
        return "│ " + " │ ".join(out) + " │"

    def sep(l, m, r, ch="─"):
        return l + m.join(ch * (w + 2) for w in width) + r
    lines = [sep("╭", "┬", "╮")]
    if titles:
        lines += [fmt(titles, center=True), sep("╞", "╪", "╡", "═")]
    for i, r in enumerate(rows):
        lines.append(fmt(r))
        if inner and i < len(rows) - 1:
            lines.append(sep("├", "┼", "┤"))
    lines.append(sep("╰", "┴", "╯"))
    return "\n".join(lines) + "\n"


def write_pretty(w, atoms, result, precision, source_name):
    p = precision
    w.write(_table([["Cheq Charge Equilibration Results"]]))
    w.write("\n")
    w.write(_table([["Source File:", source_name], ["Total Atoms:", len(atoms)],
                    ["Total Charge:", f"{sum(result.charges):.{p}f} e"],
                    ["Converged Iterations:", result.iterations],
                    ["Equilibrated Potential:", f"{result.equilibrated_potential:.{p}f} eV"]], inner=False))
    w.write("\n")
    rows = [[i, atomic_number_to_symbol(a.atomic_number), f"{a.position[0]:.{p}f}", f"{a.position[1]:.{p}f}",
             f"{a.position[2]:.{p}f}", f"{q:.{p}f}"] for i, (a, q) in enumerate(zip(atoms, result.charges))]
    w.write(_table(rows, titles=["Index", "Element", "X (Å)", "Y (Å)", "Z (Å)", "Charge (e)"],
                   align=["r", "l", "r", "r", "r", "r"]))


def write_xyz(w, atoms, result, comment, precision):
    p = precision
    w.write(f"{len(atoms)}\n")
    w.write(f"{comment.strip()} | QEq charges | iterations: {result.iterations} | potential: "
            f"{result.equilibrated_potential:.{p}f}\n")
    for a, q in zip(atoms, result.charges):
        w.write(f"{atomic_number_to_symbol(a.atomic_number)} {a.position[0]:.{p}f} {a.position[1]:.{p}f} "
                f"{a.position[2]:.{p}f} {q:.{p}f}\n")


def write_csv(w, atoms, result, precision):
    p = precision
    w.write("index,element,x,y,z,charge\n")
    for i, (a, q) in enumerate(zip(atoms, result.charges)):
        w.write(f"{i},{atomic_number_to_symbol(a.atomic_number)},{a.position[0]:.{p}f},{a.position[1]:.{p}f},"
                f"{a.position[2]:.{p}f},{q:.{p}f}\n")


def write_json(w, atoms, result, precision):
    p = precision
    w.write("{\n  \"atoms\": [\n")
    for i, (a, q) in enumerate(zip(atoms, result.charges)):
        comma = "," if i < len(atoms) - 1 else ""
        w.write("    {\n")
        w.write(f"      \"index\": {i},\n")
        w.write(f"      \"element\": \"{atomic_number_to_symbol(a.atomic_number)}\",\n")
        w.write(f"      \"position\": [{a.position[0]:.{p}f}, {a.position[1]:.{p}f}, {a.position[2]:.{p}f}],\n")
        w.write(f"      \"charge\": {q:.{p}f}\n")
        w.write(f"    }}{comma}\n")
    w.write("  ],\n")
    w.write(f"  \"total_charge\": {sum(result.charges):.{p}f},\n")
    w.write(f"  \"iterations\": {result.iterations},\n")
    w.write(f"  \"equilibrated_potential\": {result.equilibrated_potential:.{p}f}\n")
    w.write("}\n")


def write_results(w, atoms, result, comment, fmt, precision, source_name):
    if fmt == "pretty":
        write_pretty(w, atoms, result, precision, source_name)
    elif fmt == "xyz":
        write_xyz(w, atoms, result, comment, precision)
    elif fmt == "csv":
        write_csv(w, atoms, result, precision)
    else:
        write_json(w, atoms, result, precision)


# ---- argument parsing (cli.rs) ------------------------------------------------------------------------------------
def build_parser():
    ap = argparse.ArgumentParser(prog="cheq", description="A command-line tool for calculating dynamic partial atomic "
                                 "charges using the QEq method (B200 engine).")
    ap.add_argument("input", metavar="INPUT", help="XYZ file, or '-' for stdin")
    o = ap.add_argument_group("Output Options")
    o.add_argument("-o", "--output", metavar="FILE")
    o.add_argument("-f", "--format", choices=["pretty", "xyz", "csv", "json"], default="pretty")
    o.add_argument("-p", "--precision", type=int, default=6)
    c = ap.add_argument_group("Calculation Options")
    c.add_argument("-P", "--params", metavar="FILE")
    c.add_argument("-q", "--total-charge", type=float, default=0.0)
    c.add_argument("--field", type=float, nargs=3, metavar=("EX", "EY", "EZ"), help="uniform field, V/Angstrom")
    c.add_argument("--point-charges", metavar="FILE", help="lines: element x y z charge")
    c.add_argument("--trajectory", action="store_true", help="solve every frame of a multi-frame XYZ (one topology)")
    c.add_argument("--device", type=int, default=0)
    s = ap.add_argument_group("Solver Options")
    s.add_argument("--tolerance", type=float, default=1e-6)
    s.add_argument("--max-iterations", type=int, default=100)        # cli.rs:89 (library default is 2000)
    s.add_argument("--lambda-scale", type=float, default=0.5)
    s.add_argument("--hydrogen-scf", type=lambda v: v.lower() in ("true", "1", "yes"), default=True, metavar="BOOL")
    s.add_argument("--basis", choices=["sto", "gto"], default="sto")
    s.add_argument("--damping", choices=["auto", "fixed", "none"], default="auto")
    s.add_argument("--damping-factor", type=float, default=0.4)
    return ap


def options_from_args(a) -> SolverOptions:
    damping = {"auto": DampingStrategy.auto(a.damping_factor), "fixed": DampingStrategy.fixed(a.damping_factor),
               "none": DampingStrategy.none()}[a.damping]
    return SolverOptions(tolerance=a.tolerance, max_iterations=a.max_iterations, lambda_scale=a.lambda_scale,
                         hydrogen_scf=a.hydrogen_scf, basis_type=BasisType.Sto if a.basis == "sto" else BasisType.Gto,
                         damping=damping)


def run(argv=None, solver_factory=None) -> int:
    a = build_parser().parse_args(argv)
    try:
        params = Parameters.load_from_file(a.params) if a.params else get_default_parameters()
        opts = options_from_args(a)
        frames = read_frames(a.input, first_only=not a.trajectory)
        if solver_factory is None:
            from . import Context
            solver = QEqSolver(params, opts, Context(a.device))
        else:
            solver = solver_factory(params, opts)
        ext = ExternalPotential.new()
        if a.point_charges:
            ext = ext.with_point_charges(read_point_charges(a.point_charges))
        if a.field:
            ext = ext.with_uniform_field(a.field)
        out = open(a.output, "w") if a.output else sys.stdout
        try:
            if a.trajectory and len(frames) > 1 and ext.is_empty():
                import numpy as np
                from . import CalculationResult
                zs = [x.atomic_number for x in frames[0][0]]
                for atoms, _ in frames:
                    if [x.atomic_number for x in atoms] != zs:
                        raise CliError("--trajectory needs the same atoms in the same order in every frame")
                xyz = np.array([[x.position for x in atoms] for atoms, _ in frames], dtype=float)
                q, mu, its, status = solver.solve_batch(zs, xyz, a.total_charge)
                for f, (atoms, comment) in enumerate(frames):
                    if status[f] != 0:
                        raise CliError(f"frame {f}: solver status {int(status[f])}")
                    res = CalculationResult(q[f].tolist(), float(mu[f]), int(its[f]))
                    write_results(out, atoms, res, comment, a.format, a.precision, a.input)
            else:
                for atoms, comment in frames:
                    res = solver.solve_in_field(atoms, a.total_charge, ext)
                    write_results(out, atoms, res, comment, a.format, a.precision, a.input)
        finally:
            if out is not sys.stdout:
                out.close()
    except (CliError, CheqError) as e:
        sys.stderr.write(f"Error: {e}\n")
        return 1
    return 0


def main():
    sys.exit(run())
