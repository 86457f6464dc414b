#!/usr/bin/env python3
"""Arbitrary-precision ground truth for the ns-ns Slater two-centre Coulomb integral (TEST INFRASTRUCTURE).

Definition (the contract of `sto_ns::sto_coulomb_integral`, call site
/root/reference/src/shielding/sto.rs:38): with a = 2*zeta,
    rho_n(r) = a^(2n+1) / (4 pi (2n)!) * r^(2n-2) * exp(-a r)
    J_AB(R)  = int int rho_A(r1) rho_B(r2) / |r1 - r2| d3r1 d3r2          [Hartree]
evaluated here as  J = (2 pi / R) int_0^inf r rho_B(r) [G_A(R+r) - G_A(|R-r|)] dr,
G_A(s) = int_0^s t V_A(t) dt, V_A the closed-form potential of rho_A -- a direct restatement of the
definition, independent of the quadrature formula used by oracle/cheq_oracle.c and of the coefficient
tables used by the CUDA path.

Running this file regenerates tests/golden/sto_truth.json (committed).  Needs mpmath; takes ~7 min.
"""
import json
import sys
from pathlib import Path

import mpmath as mp

mp.mp.dps = 50


def G_A(n, a, s):
    """int_0^s t V_n(t) dt with V_n(t) = (1/t) [1 - exp(-a t) sum_{k<2n} (1 - k/2n) (a t)^k / k!]."""
    tot = s
    x = a * s
    ex = mp.e ** (-x)
    for k in range(2 * n):
        ck = 1 - mp.mpf(k) / (2 * n)
        inner = mp.fsum(x ** i / mp.factorial(i) for i in range(k + 1))
        tot -= ck / a * (1 - ex * inner)
    return tot


def sto_coulomb_truth(R, n1, zeta1, n2, zeta2):
    a, b, R = 2 * mp.mpf(zeta1), 2 * mp.mpf(zeta2), mp.mpf(R)
    norm = b ** (2 * n2 + 1) / (4 * mp.pi * mp.factorial(2 * n2))

    def f(r):
        return norm * r ** (2 * n2 - 1) * mp.e ** (-b * r) * (G_A(n1, a, R + r) - G_A(n1, a, abs(R - r)))

    peak = (2 * n2 - 1) / b
    pts = sorted({mp.mpf(0), R, peak, R + peak, R + 4 * peak + 10 / b, R + 10 * peak + 60 / b})
    val = mp.quad(f, pts + [mp.inf])
    return 2 * mp.pi / R * val


def zeta(n, radius_A, lam):
    """compute_slater_exponent (sto.rs:43-48) in exact-input arithmetic: the double-rounded inputs are used."""
    r_bohr = mp.mpf(radius_A) / mp.mpf(0.529177210903)
    return mp.mpf(lam) * (2 * n + 1) / (2 * r_bohr)


def main():
    import tomllib
    root = Path(__file__).resolve().parent.parent
    el = tomllib.load(open(root / "cheq_b200" / "data" / "qeq_params.toml", "rb"))["elements"]
    P = {int(k): (v["radius"], int(v["n"])) for k, v in el.items()}
    lam = 0.5
    pairs = [(1, 1), (1, 8), (8, 8), (6, 7), (7, 8), (6, 8), (1, 6), (1, 7), (3, 9), (11, 17), (55, 53),
             (19, 35), (24, 25), (5, 23), (27, 42), (32, 91), (14, 1), (15, 1), (37, 9), (9, 17), (2, 1),
             (92, 92), (87, 3)]
    Rs = ["0.05", "0.3", "0.9", "1.7", "2.5", "3.6", "5.0", "7.5", "11.0", "16.0", "25.0", "40.0", "90.0"]
    rows = []
    for (z1, z2) in pairs:
        r1, n1 = P[z1]
        r2, n2 = P[z2]
        # the double-precision zetas exactly as the reference computes them
        zd1 = float(lam * (2.0 * n1 + 1.0) / (2.0 * (r1 / 0.529177210903)))
        zd2 = float(lam * (2.0 * n2 + 1.0) / (2.0 * (r2 / 0.529177210903)))
        for Rs_ in Rs:
            Rd = float(Rs_)
            J = sto_coulomb_truth(Rd, n1, zd1, n2, zd2)
            rows.append({"z1": z1, "z2": z2, "n1": n1, "zeta1": zd1.hex(), "n2": n2, "zeta2": zd2.hex(),
                         "R_bohr": Rd.hex(), "J_hartree": mp.nstr(J, 25)})
            print(z1, z2, Rs_, mp.nstr(J, 20), file=sys.stderr)
    # synthetic near-degenerate exponents (the hard numerical regime, SURVEY.md section 7 hard part 1)
    for (n1, n2, z, rel) in [(2, 2, 0.95, 1e-2), (2, 2, 0.95, 1e-4), (2, 2, 0.95, 1e-7), (1, 2, 1.0, 3e-3),
                             (3, 3, 0.7, 5e-2), (4, 5, 0.6, 1e-3), (2, 2, 0.95, 0.12), (3, 2, 0.8, 0.2),
                             (6, 6, 0.5, 0.3), (7, 4, 0.45, 0.15)]:
        zd1, zd2 = float(z), float(z * (1 + rel))
        for Rs_ in ["0.4", "1.3", "2.9", "6.0", "12.0", "30.0"]:
            Rd = float(Rs_)
            J = sto_coulomb_truth(Rd, n1, zd1, n2, zd2)
            rows.append({"z1": 0, "z2": 0, "n1": n1, "zeta1": zd1.hex(), "n2": n2, "zeta2": zd2.hex(),
                         "R_bohr": Rd.hex(), "J_hartree": mp.nstr(J, 25)})
            print(n1, n2, z, rel, Rs_, mp.nstr(J, 20), file=sys.stderr)
    out = root / "tests" / "golden" / "sto_truth.json"
    out.write_text(json.dumps({"lambda": lam, "dps": mp.mp.dps, "rows": rows}, indent=0))
    print("wrote", out, len(rows))


if __name__ == "__main__":
    main()
