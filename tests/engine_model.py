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


def solve_model_stale(A, chi_plus_v, hard, nq, total_charge, tolerance=1e-6, max_iterations=2000,
                      damping=("auto", 0.4), freeze_ev=4.5, force_it=5, gm_tol=1e-13, trace=None):
    """Model of the stale-factor SCF iterations (cheq_b200/csrc/kernels_krylov.cu + engine.cu): after the SCF has
    settled, ONE factor of the hydrogen Schur block is kept as R = L^-T S and every later iteration solves the
    reduced bordered system  [[S0 + D_k, t1], [t1^T, -g]] [x_h; mu] = [t_c; Q - h]  by GMRES preconditioned with
    the frozen K_s^-1, restarted from the true residual.  Same policy and formulas as the engine."""
    n = A.shape[0]
    is_h = (nq == 1)
    xs = [i for i in range(n) if not is_h[i]]
    hs = [i for i in range(n) if is_h[i]]
    assert hs, "model needs an SCF set"
    nxp, nhp = round_up(len(xs)), round_up(len(hs))
    NP = nxp + nhp
    perm = -np.ones(NP, dtype=int)
    perm[:len(xs)] = xs
    perm[nxp:nxp + len(hs)] = hs
    idx = np.where(perm >= 0)[0]
    M = np.eye(NP)
    M[np.ix_(idx, idx)] = A[np.ix_(perm[idx], perm[idx])]
    c = np.zeros(NP)
    one = np.zeros(NP)
    c[idx] = -chi_plus_v[perm[idx]]
    one[idx] = 1.0
    T = np.vstack([M, c, one])
    sgn = np.ones(NP)

    def factor(T, j0, j1):
        for j in range(j0, j1):
            d = T[j, j]
            s = -1.0 if d < 0 else 1.0
            sgn[j] = s
            l = np.sqrt(abs(d))
            T[j + 1:, j] *= s / l
            T[j, j] = l
            T[j + 1:, j + 1:j1] -= s * np.outer(T[j + 1:, j], T[j + 1:j1, j])
        return T

    T = factor(T, 0, nxp)
    T[nxp:, nxp:] -= (T[nxp:, :nxp] * sgn[:nxp]) @ T[nxp:NP, :nxp].T
    S0 = T[nxp:, nxp:].copy()                       # (nhp + 2) x nhp: Schur block + reduced right-hand sides
    S0sym = np.tril(S0[:nhp]) + np.tril(S0[:nhp], -1).T
    tc, t1 = S0[nhp].copy(), S0[nhp + 1].copy()
    y1x, ycx = T[NP + 1, :nxp], T[NP, :nxp]
    g = float((sgn[:nxp] * y1x) @ y1x)
    h = float((sgn[:nxp] * y1x) @ ycx)
    b = np.concatenate([tc, [total_charge - h]])
    hard_p = np.zeros(NP)
    hard_p[idx] = hard[perm[idx]]
    real_h = perm[nxp:] >= 0

    def shift(q):
        d = np.zeros(nhp)
        d[real_h] = hard_p[nxp:][real_h] * (np.clip(q[nxp:][real_h], -0.95, 0.95) * H_FACTOR)
        return d

    def Kmul(d, v):
        return np.concatenate([S0sym @ v[:nhp] + d * v[:nhp] + v[nhp] * t1, [t1 @ v[:nhp] - g * v[nhp]]])

    frozen = {}

    def precond(r):
        R, sh, ws, den = frozen["R"], frozen["sgn"], frozen["ws"], frozen["den"]
        u0 = R @ (sh * (R.T @ r[:nhp]))
        nu = (t1 @ u0 - r[nhp]) / den
        return np.concatenate([u0 - nu * ws, [nu]])

    def gmres(d, sol):
        steps, scale, beta_prev = 0, 0.0, 0.0
        m = 48
        for cycle in range(8):
            z = precond(b - Kmul(d, sol))
            steps += 1
            beta = np.linalg.norm(z)
            if cycle == 0:
                scale = max(np.linalg.norm(sol), 1e-300)
            if beta <= gm_tol * scale:
                return sol, steps, True
            if cycle > 0 and beta > 0.25 * beta_prev:
                return sol, steps, beta <= 1e3 * gm_tol * scale
            beta_prev = beta
            V = [z / beta]
            H = np.zeros((m + 1, m))
            gv = np.zeros(m + 1)
            gv[0] = beta
            cs, sn = [], []
            k = 0
            for j in range(m):
                w = precond(Kmul(d, V[j]))
                steps += 1
                hsum = np.zeros(j + 1)
                for _ in range(2):
                    hh = np.array([V[i] @ w for i in range(j + 1)])
                    w = w - sum(hh[i] * V[i] for i in range(j + 1))
                    hsum += hh
                hn = np.linalg.norm(w)
                V.append(w / hn if hn > 0 else w * 0)
                H[:j + 1, j] = hsum
                for i in range(j):
                    a = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                    H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                    H[i, j] = a
                rr = np.hypot(H[j, j], hn)
                cs.append(H[j, j] / rr)
                sn.append(hn / rr)
                H[j, j] = rr
                gv[j + 1] = -sn[j] * gv[j]
                gv[j] = cs[j] * gv[j]
                k = j + 1
                if abs(gv[j + 1]) <= 0.3 * gm_tol * scale:
                    break
            y = np.linalg.solve(np.triu(H[:k, :k]), gv[:k])
            sol = sol + sum(y[i] * V[i] for i in range(k))
        return sol, steps, False

    q = np.zeros(NP)
    d = 1.0 if damping[0] == "none" else damping[1]
    prev = np.finfo(float).max
    d_prev = np.zeros(nhp)
    have_factor = False
    sol = np.zeros(nhp + 1)
    for it in range(1, max_iterations + 1):
        d_cur = shift(q)
        use_krylov = bool(frozen)
        if not frozen and have_factor and it >= 2:
            if np.max(np.abs(d_cur - d_prev)) <= freeze_ev or it >= force_it:
                Lh = np.tril(T[nxp:NP, nxp:])
                sh = sgn[nxp:].copy()
                R = np.linalg.solve(Lh, np.eye(nhp)).T * sh[None, :]      # L^-T S
                ws = R @ (sh * (R.T @ t1))
                frozen.update(R=R, sgn=sh, ws=ws, den=float(t1 @ ws + g))
                use_krylov = True
        steps = 0
        if use_krylov:
            sol, steps, ok = gmres(d_cur, sol)
            assert ok, "GMRES on the frozen factor did not converge"
            mu = sol[nhp]
            x = np.zeros(NP)
            x[nxp:] = sol[:nhp]
            v = ycx - mu * y1x
            Lxx = np.tril(T[:nxp, :nxp])
            W = T[nxp:NP, :nxp]                         # L[h rows, x cols]
            x[:nxp] = np.linalg.solve(Lxx.T, v - W.T @ x[nxp:])
        else:
            T[nxp:, nxp:] = S0
            hh = np.arange(nhp)
            T[nxp + hh, nxp + hh] = S0[hh, hh] + d_cur
            T = factor(T, nxp, NP)
            fc, f1 = T[NP, :], T[NP + 1, :]
            mu = ((sgn * f1) @ fc - total_charge) / ((sgn * f1) @ f1)
            x = np.linalg.solve(np.tril(T[:NP, :]).T, fc - mu * f1)
            d_prev = d_cur
            have_factor = True
            sol = np.concatenate([x[nxp:], [mu]])
        if trace is not None:
            trace.append((1 if use_krylov else 0, steps))
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
