"""NumPy model of the engine's DISTRIBUTED schedule (TEST INFRASTRUCTURE; mirrors factor_blocked / factor_and_scf in
cheq_b200/csrc/engine.cu): 1-D block-cyclic column panels with the owner map the C library computes
(cheq_dist_plan), owner-computes assembly and trailing updates, one broadcast per factored panel, replicated
triangular solves and SCF control.  Columns a rank does not own start as NaN and only ever become valid by
receiving a broadcast panel, so a schedule that reads a column before it has arrived poisons the result.

`bcast(array, root)` is the only communication primitive (torch.distributed.broadcast over gloo in the world-size-2
CPU test, identity for one rank)."""
import numpy as np

H_FACTOR = 0.93475415965


def dist_solve_model(A, chi_plus_v, hard, nq, total_charge, rank, world, bcast, owner, panel, tile,
                     tolerance=1e-6, max_iterations=200, damping=0.4):
    n = A.shape[0]
    is_h = (nq == 1)
    xs = [i for i in range(n) if not is_h[i]]
    hs = [i for i in range(n) if is_h[i]]
    up = lambda v: (v + tile - 1) // tile * tile
    nxp, nhp = up(len(xs)), up(len(hs))
    NP = nxp + nhp
    assert len(owner) == NP // tile
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
    col_owner = np.repeat(owner, tile)
    mine = col_owner == rank
    T[:, ~mine] = np.nan                      # owner-computes assembly: other ranks' columns do not exist here
    T[np.triu_indices(NP, 1)] = np.nan        # only the lower triangle is ever stored
    sgn = np.full(NP, np.nan)
    panels = []
    for p in range(panel.max() + 1):
        t = np.where(panel == p)[0]
        panels.append((t[0] * tile, (t[-1] + 1) * tile, int(owner[t[0]])))
    x_panels = sum(1 for (a, b, o) in panels if b <= nxp)

    def factor_blocked(p0, p1):
        for (j0, j1, root) in panels[p0:p1]:
            if root == rank:
                for j in range(j0, j1):
                    d = T[j, j]
                    assert abs(d) > 1e-4
                    s = -1.0 if d < 0 else 1.0
                    sgn[j] = s
                    l = np.sqrt(abs(d))
                    T[j + 1:, j] *= s / l
                    T[j, j] = l
                    for k in range(j + 1, j1):
                        T[k:, k] -= s * T[k:, j] * T[k, j]
            buf = np.ascontiguousarray(T[j0:, j0:j1]) if root == rank else np.empty((NP + 2 - j0, j1 - j0))
            bcast(buf, root)
            sg = np.ascontiguousarray(sgn[j0:j1]) if root == rank else np.empty(j1 - j0)
            bcast(sg, root)
            if root != rank:
                T[j0:, j0:j1] = buf
                sgn[j0:j1] = sg
            W = T[:, j0:j1]
            for k in np.where(mine)[0]:
                if k >= j1:
                    T[k:, k] -= (W[k:, :] * sgn[j0:j1]) @ W[k, :]

    def finish():
        fc, f1 = T[NP, :], T[NP + 1, :]
        mu = ((sgn * f1) @ fc - total_charge) / ((sgn * f1) @ f1)
        v = fc - mu * f1
        L = np.tril(T[:NP, :])
        return np.linalg.solve(L.T, v), mu

    factor_blocked(0, x_panels)
    own_h = np.where(mine & (np.arange(NP) >= nxp))[0]
    S0 = T[:, own_h].copy()                   # snapshot of the owned hydrogen columns (Schur complement + rhs rows)
    hard_p = np.zeros(NP)
    hard_p[idx] = hard[perm[idx]]
    q = np.zeros(NP)
    d = damping
    prev = np.finfo(float).max
    real_h = [k for k in own_h if perm[k] >= 0]
    for it in range(1, max_iterations + 1):
        T[:, own_h] = S0
        for k in real_h:
            T[k, k] = S0[k, list(own_h).index(k)] + hard_p[k] * (np.clip(q[k], -0.95, 0.95) * H_FACTOR)
        factor_blocked(x_panels, len(panels))
        x, mu = finish()
        delta = np.max(np.abs(x[idx] - q[idx]))
        q[idx] = (1.0 - d) * q[idx] + d * x[idx]
        state = np.array([delta, mu, d])
        bcast(state, 0)                       # rank 0's decision is authoritative
        bcast(q, 0)
        delta, mu = state[0], state[1]
        if delta < tolerance:
            out = np.zeros(n)
            out[perm[idx]] = q[idx]
            return out, mu, it
        if delta > prev:
            d *= 0.5
        elif delta < prev * 0.9:
            d = min(d * 1.1, 1.0)
        d = max(d, 0.001)
        prev = delta
    raise RuntimeError("not converged")
