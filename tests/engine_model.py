"""NumPy model of the engine's algebra (TEST INFRASTRUCTURE): permutation + identity padding to 128-tiles,
trapezoid Cholesky with the two right-hand sides as extra rows, hydrogen Schur complement reused across SCF
iterations, constraint elimination, and the reference's damping/convergence logic.  It exists to show on the CPU
that the reformulation the CUDA path uses reproduces the bordered partial-pivot LU of
/root/reference/src/solver/implementation.rs:557-603 to rounding; it is not part of the product."""
import numpy as np

H_FACTOR = 0.93475415965
TILE = 128


def round_up(v):
    return (v + TILE - 1) // TILE * TILE


def solve_model(A, chi_plus_v, hard, nq, total_charge, tolerance=1e-6, max_iterations=2000, hydrogen_scf=True,
                damping=("auto", 0.4)):
    """A: n x n symmetric (hardness diagonal + J).  Returns charges, mu, iterations."""
    n = A.shape[0]
    is_h = (nq == 1)
    scf = hydrogen_scf and is_h.any()
    xs = [i for i in range(n) if not (scf and is_h[i])]
    hs = [i for i in range(n) if scf and is_h[i]]
    nxp, nhp = round_up(len(xs)), round_up(len(hs))
    NP = nxp + nhp
    perm = -np.ones(NP, dtype=int)
    perm[:len(xs)] = xs
    perm[nxp:nxp + len(hs)] = hs
    real = perm >= 0
    M = np.eye(NP)
    idx = np.where(real)[0]
    M[np.ix_(idx, idx)] = A[np.ix_(perm[idx], perm[idx])]
    c = np.zeros(NP)
    one = np.zeros(NP)
    c[idx] = -chi_plus_v[perm[idx]]
    one[idx] = 1.0
    T = np.vstack([M, c, one])  # trapezoid (NP + 2) x NP

    sgn = np.ones(NP)

    def factor(T, j0, j1):
        """in-place A = L S L^T (S = +-1, no pivoting) of columns [j0, j1) of the trapezoid, rows below included"""
        for j in range(j0, j1):
            d = T[j, j]
            if not abs(d) > 1e-4:
                raise np.linalg.LinAlgError("vanishing pivot")
            s = -1.0 if d < 0 else 1.0
            sgn[j] = s
            l = np.sqrt(abs(d))
            T[j + 1:, j] *= s / l
            T[j, j] = l
            T[j + 1:, j + 1:j1] -= s * np.outer(T[j + 1:, j], T[j + 1:j1, j])
        return T

    def finish(T):
        fc, f1 = T[NP, :], T[NP + 1, :]           # y = S L^-1 b
        mu = ((sgn * f1) @ fc - total_charge) / ((sgn * f1) @ f1)
        v = fc - mu * f1
        L = np.tril(T[:NP, :])
        x = np.linalg.solve(L.T, v)
        return x, mu

    q = np.zeros(NP)
    if not scf:
        T = factor(T, 0, NP)
        x, mu = finish(T)
        out = np.zeros(n)
        out[perm[idx]] = x[idx]
        return out, mu, 1
    T = factor(T, 0, nxp)
    T[nxp:, nxp:] -= (T[nxp:, :nxp] * sgn[:nxp]) @ T[nxp:NP, :nxp].T
    S0 = T[nxp:, nxp:].copy()
    hard_p = np.zeros(NP)
    hard_p[idx] = hard[perm[idx]]
    d = 1.0 if damping[0] == "none" else damping[1]
    prev = np.finfo(float).max
    for it in range(1, max_iterations + 1):
        T[nxp:, nxp:] = S0
        hh = np.arange(nxp, nxp + len(hs))
        qc = np.clip(q[hh], -0.95, 0.95)
        T[hh, hh] = S0[hh - nxp, hh - nxp] + hard_p[hh] * (qc * H_FACTOR)
        T = factor(T, nxp, NP)
        x, mu = finish(T)
        delta = np.max(np.abs(x[idx] - q[idx]))
        q[idx] = (1.0 - d) * q[idx] + d * x[idx]
        if delta < tolerance:
            out = np.zeros(n)
            out[perm[idx]] = q[idx]
            return out, mu, it
        if damping[0] == "auto":
            if delta > prev:
                d *= 0.5
            elif delta < prev * 0.9:
                d = min(d * 1.1, 1.0)
            d = max(d, 0.001)
        prev = delta
    raise RuntimeError("not converged")
