// See sto_tables.h.  Pure host code (g++), long double throughout the construction.
//
// Mathematics (derived in DESIGN.md section 3; checked against 40-digit values in tests/golden/sto_truth.json):
// with a = 2 zeta1, b = 2 zeta2, al = 2 n1 - 1, be = 2 n2 - 1, r = 1 + i + j, c(x) = b + x (a - b),
//   F(R) = 1 - R J(R)
//        = N sum_{i<=al, j<=be} C(al,i) C(be,j) (al+be-i-j)! / (a+b)^(1+al+be-i-j)
//            int_0^1 x^i (1-x)^j exp(-c R) sum_{l<=r} C(r,l) (l+1)! R^(r-l) c^(-2-l) dx,
//   N = a^(al+2) b^(be+2) / ((al+1)! (be+1)!).
// All terms are positive.  Three evaluation forms are derived from it:
//   NODES      Gauss-Legendre in x: F = sum_g exp(-c_g R) P_g(R), always valid, one exp per node;
//   COLLAPSED  exp(-c_g R) = exp(-cbar R) exp(-(c_g - cbar) R), the second factor Taylor-expanded to order
//              M and folded into ONE polynomial: F = exp(-cbar R) S(R).  Exact for equal exponents (M = 0)
//              and uniformly stable for near-degenerate ones, where the textbook form cancels catastrophically;
//   CLOSED     the textbook partial-fraction form F = exp(-aR) P(R) + exp(-bR) Q(R), used only when its
//              measured fp64 error is within tolerance (well separated exponents).
#include "sto_tables.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <mutex>

namespace cheq {
namespace {

using ld = long double;
constexpr int kMaxN = 15;          // principal quantum numbers 1..15 (table uses 1..7)
constexpr int kMaxR = 2 * kMaxN * 2 + 2;

struct GaussLegendre {
    std::vector<ld> x, w;  // nodes and weights on [0, 1]
};

const GaussLegendre& gauss_legendre(int n) {
    static std::map<int, GaussLegendre> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    GaussLegendre g;
    g.x.resize(n);
    g.w.resize(n);
    const ld pi = acosl(-1.0L);
    for (int k = 0; k < n; ++k) {
        ld t = cosl(pi * (k + 0.75L) / (n + 0.5L));
        ld dp = 1.0L;
        for (int it2 = 0; it2 < 100; ++it2) {
            ld p0 = 1.0L, p1 = t;
            for (int m = 2; m <= n; ++m) {
                ld p2 = ((2 * m - 1) * t * p1 - (m - 1) * p0) / m;
                p0 = p1;
                p1 = p2;
            }
            dp = n * (t * p1 - p0) / (t * t - 1.0L);
            ld dt = p1 / dp;
            t -= dt;
            if (fabsl(dt) < 1e-19L) break;
        }
        ld p0 = 1.0L, p1 = t;
        for (int m = 2; m <= n; ++m) {
            ld p2 = ((2 * m - 1) * t * p1 - (m - 1) * p0) / m;
            p0 = p1;
            p1 = p2;
        }
        dp = n * (t * p1 - p0) / (t * t - 1.0L);
        g.x[k] = 0.5L * (1.0L - t);
        g.w[k] = 1.0L / ((1.0L - t * t) * dp * dp);
    }
    return cache.emplace(n, std::move(g)).first->second;
}

struct Combinatorics {
    ld fact[96];
    ld binom[96][96];
    Combinatorics() {
        fact[0] = 1.0L;
        for (int k = 1; k < 96; ++k) fact[k] = fact[k - 1] * k;
        for (int r = 0; r < 96; ++r)
            for (int l = 0; l < 96; ++l) binom[r][l] = 0.0L;
        for (int r = 0; r < 96; ++r) {
            binom[r][0] = 1.0L;
            for (int l = 1; l <= r; ++l) binom[r][l] = binom[r - 1][l - 1] + binom[r - 1][l];
        }
    }
};
const Combinatorics& comb() {
    static Combinatorics c;
    return c;
}

struct PairSetup {
    int al, be, nr;  // nr = al + be + 2 = number of polynomial coefficients per node
    ld a, b, N;
    ld w[2 * kMaxN][2 * kMaxN];
};

PairSetup make_setup(int n1, double zeta1, int n2, double zeta2) {
    const Combinatorics& C = comb();
    PairSetup s;
    s.al = 2 * n1 - 1;
    s.be = 2 * n2 - 1;
    s.nr = s.al + s.be + 2;
    s.a = 2.0L * (ld)zeta1;
    s.b = 2.0L * (ld)zeta2;
    s.N = powl(s.a, s.al + 2) * powl(s.b, s.be + 2) / (C.fact[s.al + 1] * C.fact[s.be + 1]);
    const ld ab = s.a + s.b;
    for (int i = 0; i <= s.al; ++i)
        for (int j = 0; j <= s.be; ++j) {
            const int m = s.al + s.be - i - j;
            s.w[i][j] = C.binom[s.al][i] * C.binom[s.be][j] * C.fact[m] / powl(ab, 1 + m);
        }
    return s;
}

// Polynomial (in R) carried by quadrature abscissa x, *without* the weight and the factor N.
void node_polynomial(const PairSetup& s, ld x, ld* poly, ld* c_out) {
    const Combinatorics& C = comb();
    const ld omx = 1.0L - x;
    const ld c = s.b + x * (s.a - s.b);
    const ld ic = 1.0L / c;
    ld xp[2 * kMaxN], yp[2 * kMaxN], icp[kMaxR + 4];
    xp[0] = yp[0] = 1.0L;
    for (int i = 1; i <= s.al; ++i) xp[i] = xp[i - 1] * x;
    for (int j = 1; j <= s.be; ++j) yp[j] = yp[j - 1] * omx;
    icp[0] = 1.0L;
    for (int k = 1; k <= s.nr + 2; ++k) icp[k] = icp[k - 1] * ic;
    for (int k = 0; k < s.nr; ++k) poly[k] = 0.0L;
    for (int r = 1; r < s.nr; ++r) {
        ld Wr = 0.0L;
        for (int i = 0; i <= s.al; ++i) {
            const int j = r - 1 - i;
            if (j < 0 || j > s.be) continue;
            Wr += s.w[i][j] * xp[i] * yp[j];
        }
        for (int l = 0; l <= r; ++l) poly[r - l] += Wr * C.binom[r][l] * C.fact[l + 1] * icp[2 + l];
    }
    *c_out = c;
}

struct NodeSet {
    int G = 0;
    std::vector<ld> c, wN;    // exponent and weight * N per node
    std::vector<ld> poly;     // G x nr
};

NodeSet make_nodes(const PairSetup& s, int G) {
    const GaussLegendre& gl = gauss_legendre(G);
    NodeSet ns;
    ns.G = G;
    ns.c.resize(G);
    ns.wN.resize(G);
    ns.poly.resize((size_t)G * s.nr);
    for (int g = 0; g < G; ++g) {
        node_polynomial(s, gl.x[g], &ns.poly[(size_t)g * s.nr], &ns.c[g]);
        ns.wN[g] = gl.w[g] * s.N;
    }
    return ns;
}

ld eval_nodes(const PairSetup& s, const NodeSet& ns, ld R) {
    ld total = 0.0L;
    for (int g = 0; g < ns.G; ++g) {
        const ld* p = &ns.poly[(size_t)g * s.nr];
        ld acc = 0.0L;
        for (int k = s.nr - 1; k >= 0; --k) acc = acc * R + p[k];
        total += ns.wN[g] * acc * expl(-ns.c[g] * R);
    }
    return total;
}

ld j0_nodes(const PairSetup& s, const NodeSet& ns) {
    ld total = 0.0L;  // J(0) = -F'(0) = sum_g w_g (c_g p_g0 - p_g1)
    for (int g = 0; g < ns.G; ++g) {
        const ld* p = &ns.poly[(size_t)g * s.nr];
        total += ns.wN[g] * (ns.c[g] * p[0] - p[1]);
    }
    return total;
}

// Adaptive reference: doubles the Gauss-Legendre order until two successive orders agree.
struct Reference {
    PairSetup s;
    std::vector<NodeSet> levels;  // orders 32, 64, 128, 256, 512
    explicit Reference(const PairSetup& setup) : s(setup) {}
    const NodeSet& level(size_t k) {
        while (levels.size() <= k) levels.push_back(make_nodes(s, 32 << levels.size()));
        return levels[k];
    }
    ld F(ld R) {
        ld prev = eval_nodes(s, level(0), R);
        for (size_t k = 1; k < 5; ++k) {
            ld cur = eval_nodes(s, level(k), R);
            bool done = fabsl(cur - prev) <= 2e-19L + 2e-18L * fabsl(cur);
            prev = cur;
            if (done) break;
        }
        return prev;
    }
    ld J0() {
        ld prev = j0_nodes(s, level(0));
        for (size_t k = 1; k < 5; ++k) {
            ld cur = j0_nodes(s, level(k));
            bool done = fabsl(cur - prev) <= 4e-18L * fabsl(cur);
            prev = cur;
            if (done) break;
        }
        return prev;
    }
};

double eval_terms_fp64(const std::vector<StoTerm>& terms, double R) {
    double F = 0.0;
    for (const StoTerm& t : terms) {
        double acc = 0.0;
        for (int k = (int)t.coef.size() - 1; k >= 0; --k) acc = std::fma(acc, R, t.coef[k]);
        F = std::fma(std::exp(-t.c * R), acc, F);
    }
    return F;
}

double max_error(const std::vector<StoTerm>& terms, const std::vector<double>& grid,
                 const std::vector<ld>& truth) {
    double worst = 0.0;
    for (size_t k = 0; k < grid.size(); ++k) {
        double err = (double)fabsl((ld)eval_terms_fp64(terms, grid[k]) - truth[k]);
        if (!(err <= worst)) worst = err;  // also catches NaN
        if (!std::isfinite(err)) return std::numeric_limits<double>::infinity();
    }
    return worst;
}

int terms_cost(const std::vector<StoTerm>& terms) {
    int c = 0;
    for (const StoTerm& t : terms) c += 24 + (int)t.coef.size();
    return c;
}

// COLLAPSED form of Taylor order M built from a fine node set.
std::vector<StoTerm> build_collapsed(const PairSetup& s, const NodeSet& ns, int M) {
    const Combinatorics& C = comb();
    const ld cbar = 0.5L * (s.a + s.b);
    std::vector<ld> S((size_t)s.nr + M, 0.0L);
    for (int g = 0; g < ns.G; ++g) {
        const ld* p = &ns.poly[(size_t)g * s.nr];
        const ld d = ns.c[g] - cbar;
        ld e = 1.0L;  // (-d)^m / m!
        for (int m = 0; m <= M; ++m) {
            if (m > 0) e *= -d / (ld)m;
            for (int k = 0; k < s.nr; ++k) S[(size_t)k + m] += ns.wN[g] * p[k] * e;
        }
    }
    (void)C;
    StoTerm t;
    t.c = (double)cbar;
    t.coef.resize(S.size());
    // exp(-cbar R) is evaluated with the fp64-rounded exponent; fold the rounding into nothing: the
    // difference exp(-(cbar - fl(cbar)) R) - 1 is < 1e-16 * cbar * R * ... and is covered by verification.
    for (size_t k = 0; k < S.size(); ++k) t.coef[k] = (double)S[k];
    while (t.coef.size() > 1 && t.coef.back() == 0.0) t.coef.pop_back();
    return {t};
}

// NODES form.
std::vector<StoTerm> build_nodes(const PairSetup& s, int G) {
    NodeSet ns = make_nodes(s, G);
    std::vector<StoTerm> terms(G);
    for (int g = 0; g < G; ++g) {
        terms[g].c = (double)ns.c[g];
        terms[g].coef.resize(s.nr);
        for (int k = 0; k < s.nr; ++k) terms[g].coef[k] = (double)(ns.wN[g] * ns.poly[(size_t)g * s.nr + k]);
    }
    return terms;
}

// CLOSED form by partial fractions of the Fourier-space product (a != b).
//   rho~_n(k) = sum_{t<n} A_j / (k^2 + a^2)^j, j = 2n - t, A_j = (-1)^t a^(2j) S_{n,t} / (2n),
//   S_{n,t} = sum_{p=t}^{n-1} C(2n, 2p+1) C(p, t);
//   (2/pi) int_0^inf sin(kR) / (k (k^2+a^2)^m) dk = a^(-2m) [1 - exp(-aR) p_m(aR)],
//   p_1 = 1, p_{m+1}(x) = p_m + x/(2m) (p_m - p_m').
std::vector<StoTerm> build_closed(int n1, int n2, const PairSetup& s) {
    const Combinatorics& C = comb();
    const ld a = s.a, b = s.b;
    const ld D = b * b - a * a;
    const int mmax = 2 * std::max(n1, n2);
    std::vector<std::vector<ld>> pm(mmax + 1);
    pm[1] = {1.0L};
    for (int m = 1; m < mmax; ++m) {
        const std::vector<ld>& p = pm[m];
        std::vector<ld> q(p.size());
        for (size_t i = 0; i < p.size(); ++i) q[i] = p[i] - (i + 1 < p.size() ? p[i + 1] * (ld)(i + 1) : 0.0L);
        std::vector<ld> nxt(p.size() + 1, 0.0L);
        for (size_t i = 0; i < p.size(); ++i) nxt[i] = p[i];
        for (size_t i = 0; i < q.size(); ++i) nxt[i + 1] += q[i] / (2.0L * m);
        pm[m + 1] = nxt;
    }
    auto density_coefs = [&](int n, ld e) {
        std::vector<ld> A(2 * n + 1, 0.0L);
        for (int t = 0; t < n; ++t) {
            ld S = 0.0L;
            for (int p = t; p < n; ++p) S += C.binom[2 * n][2 * p + 1] * C.binom[p][t];
            const int j = 2 * n - t;
            A[j] = ((t & 1) ? -1.0L : 1.0L) * powl(e, 2 * j) * S / (2.0L * n);
        }
        return A;
    };
    std::vector<ld> A = density_coefs(n1, a), B = density_coefs(n2, b);
    std::vector<ld> P(2 * n1, 0.0L), Q(2 * n2, 0.0L);
    for (int j = n1 + 1; j <= 2 * n1; ++j)
        for (int jp = n2 + 1; jp <= 2 * n2; ++jp) {
            const ld AB = A[j] * B[jp];
            for (int m = 1; m <= j; ++m) {
                ld alpha = (((j - m) & 1) ? -1.0L : 1.0L) * C.binom[jp + j - m - 1][j - m] / powl(D, jp + j - m);
                ld cf = AB * alpha / powl(a, 2 * m);
                ld ap = 1.0L;
                for (size_t i = 0; i < pm[m].size(); ++i) {
                    P[i] += cf * pm[m][i] * ap;
                    ap *= a;
                }
            }
            for (int m = 1; m <= jp; ++m) {
                ld beta = (((jp - m) & 1) ? -1.0L : 1.0L) * C.binom[j + jp - m - 1][jp - m] / powl(-D, j + jp - m);
                ld cf = AB * beta / powl(b, 2 * m);
                ld bp = 1.0L;
                for (size_t i = 0; i < pm[m].size(); ++i) {
                    Q[i] += cf * pm[m][i] * bp;
                    bp *= b;
                }
            }
        }
    std::vector<StoTerm> terms(2);
    terms[0].c = (double)a;
    terms[1].c = (double)b;
    for (ld v : P) terms[0].coef.push_back((double)v);
    for (ld v : Q) terms[1].coef.push_back((double)v);
    return terms;
}

}  // namespace

int StoPairTable::cost() const { return terms_cost(terms); }

long double sto_coulomb_quadrature(int n1, double zeta1, int n2, double zeta2, long double R) {
    if (n1 < 1 || n2 < 1 || n1 > kMaxN || n2 > kMaxN || !(zeta1 > 0.0) || !(zeta2 > 0.0))
        return std::numeric_limits<long double>::quiet_NaN();
    Reference ref(make_setup(n1, zeta1, n2, zeta2));
    const ld cbar = 0.5L * (ref.s.a + ref.s.b);
    if (R * cbar < 1e-7L) return ref.J0();
    return (1.0L - ref.F(R)) / R;
}

StoPairTable build_sto_pair_table(int n1, double zeta1, int n2, double zeta2, double tol) {
    StoPairTable tab;
    tab.n1 = n1;
    tab.n2 = n2;
    tab.zeta1 = zeta1;
    tab.zeta2 = zeta2;
    if (n1 < 1 || n2 < 1 || n1 > kMaxN || n2 > kMaxN || !(zeta1 > 0.0) || !(zeta2 > 0.0) ||
        !std::isfinite(zeta1) || !std::isfinite(zeta2)) {
        tab.kind = STO_KIND_INVALID;
        tab.j0 = std::numeric_limits<double>::quiet_NaN();
        return tab;
    }
    PairSetup s = make_setup(n1, zeta1, n2, zeta2);
    Reference ref(s);
    const ld cbar = 0.5L * (s.a + s.b);
    const ld cmin = std::min(s.a, s.b);

    // range: F < 1e-19 beyond r_cut
    ld r_end = 1.0L / cmin;
    for (int it = 0; it < 400 && ref.F(r_end) > 1e-19L; ++it) r_end *= 1.15L;
    tab.r_cut = (double)r_end;
    tab.r_zero = (double)(1e-7L / cbar);
    tab.j0 = (double)ref.J0();

    // verification grid, denser at short range where F is large
    const int K = 72;
    std::vector<double> grid(K);
    std::vector<ld> truth(K);
    for (int k = 0; k < K; ++k) {
        const double u = (k + 1.0) / K;
        grid[k] = (double)r_end * u * u;
        truth[k] = ref.F((ld)grid[k]);
    }

    struct Candidate {
        std::vector<StoTerm> terms;
        int kind = STO_KIND_INVALID;
        int M = 0;
        double err = std::numeric_limits<double>::infinity();
    };
    Candidate best;        // cheapest candidate within tolerance
    Candidate fallback;    // most accurate candidate seen
    auto consider = [&](std::vector<StoTerm>&& terms, int kind, int M) {
        double err = max_error(terms, grid, truth);
        if (err < fallback.err) {
            fallback.terms = terms;
            fallback.kind = kind;
            fallback.M = M;
            fallback.err = err;
        }
        if (err <= tol && (best.kind == STO_KIND_INVALID || terms_cost(terms) < terms_cost(best.terms))) {
            best.terms = std::move(terms);
            best.kind = kind;
            best.M = M;
            best.err = err;
            return true;
        }
        return err <= tol;
    };

    // COLLAPSED
    {
        const NodeSet& fine = ref.level(2);  // 128 nodes: exact for the polynomial part up to degree 255
        if (s.a == s.b) {
            consider(build_collapsed(s, fine, 0), STO_KIND_COLLAPSED, 0);
        } else {
            static const int orders[] = {4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64};
            for (int M : orders)
                if (consider(build_collapsed(s, fine, M), STO_KIND_COLLAPSED, M)) break;
        }
    }
    // CLOSED
    if (s.a != s.b) consider(build_closed(n1, n2, s), STO_KIND_CLOSED, 0);
    // NODES (only if nothing cheaper is within tolerance)
    if (best.kind == STO_KIND_INVALID) {
        static const int orders[] = {2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128};
        for (int G : orders)
            if (consider(build_nodes(s, G), STO_KIND_NODES, 0)) break;
    }
    const Candidate& pick = (best.kind != STO_KIND_INVALID) ? best : fallback;
    tab.terms = pick.terms;
    tab.kind = pick.kind;
    tab.taylor_order = pick.M;
    tab.max_abs_err = pick.err;
    return tab;
}

double eval_sto_pair_table(const StoPairTable& t, double R) {
    if (t.kind == STO_KIND_INVALID) return std::numeric_limits<double>::quiet_NaN();
    if (R < t.r_zero) return t.j0;
    if (R > t.r_cut) return 1.0 / R;
    return (1.0 - eval_terms_fp64(t.terms, R)) / R;
}

}  // namespace cheq
