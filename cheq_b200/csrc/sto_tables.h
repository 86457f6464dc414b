// Host-side construction of per-pair-type evaluation tables for the exact ns-ns Slater Coulomb integral.
//
// Replaces `sto_ns::sto_coulomb_integral` (un-vendored crate sto-ns ^0.1.1; call site
// /root/reference/src/shielding/sto.rs:38).  For one pair type (n1, zeta1, n2, zeta2) the integral is
//     J(R) = (1 - F(R)) / R,     F(R) = sum_{g < nterms} exp(-c_g R) * P_g(R)
// with P_g polynomials whose coefficients are produced here in long double and verified, as evaluated in
// fp64, against an all-positive quadrature of the definition.  The CUDA assembly kernels
// only ever evaluate exp + Horner; all numerically delicate work (near-degenerate exponents) happens here,
// once per distinct pair of element types in a system.
#pragma once
#include <cstdint>
#include <vector>

namespace cheq {

enum StoTableKind : int32_t {
    STO_KIND_COLLAPSED = 0,  // one exponential at the mean exponent, Taylor-collapsed in the exponent gap
    STO_KIND_CLOSED = 1,     // textbook two-exponential closed form (well separated exponents only)
    STO_KIND_NODES = 2,      // Gauss-Legendre sum of all-positive exponential terms (always valid)
    STO_KIND_INVALID = 3,    // n == 0 or zeta <= 0: the integral is undefined (NaN, like the oracle)
};

struct StoTerm {
    double c;                  // exponent [1/bohr]
    std::vector<double> coef;  // P_g(R) = sum_k coef[k] R^k, R in bohr
};

struct StoPairTable {
    int32_t n1 = 0, n2 = 0;
    double zeta1 = 0, zeta2 = 0;
    int32_t kind = STO_KIND_INVALID;
    int32_t taylor_order = 0;   // M for STO_KIND_COLLAPSED
    std::vector<StoTerm> terms;
    double j0 = 0;              // J(R -> 0) [Hartree]
    double r_zero = 0;          // below this R [bohr] use j0
    double r_cut = 0;           // beyond this R [bohr] F(R) < 1e-19: J = 1/R to fp64
    double max_abs_err = 0;     // max |F_fp64(table) - F_true| over the verification grid
    int cost() const;           // rough per-pair cost in FMA-equivalents (exp counted as 24)
};

// Builds (and verifies) the table.  tol is the accepted absolute error of F as evaluated in fp64.
StoPairTable build_sto_pair_table(int n1, double zeta1, int n2, double zeta2, double tol = 1.0e-15);

// fp64 evaluation of a table exactly as the device does it (exp + Horner); Hartree, R in bohr.
double eval_sto_pair_table(const StoPairTable& t, double R_bohr);

// Reference value by adaptive all-positive quadrature in long double (used for verification).
long double sto_coulomb_quadrature(int n1, double zeta1, int n2, double zeta2, long double R_bohr);

}  // namespace cheq
