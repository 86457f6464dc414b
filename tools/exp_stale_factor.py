#!/usr/bin/env python3
"""CPU experiment (numpy/scipy, uses the oracle's assembly): how many preconditioned GMRES steps does the
reduced bordered hydrogen system need per SCF iteration when the preconditioner is the factor of an EARLIER
iteration?  Calibrates the refactor policy of the engine (stale-factor refinement)."""
import sys, time, json
from pathlib import Path
import numpy as np
from scipy.linalg import lu_factor, lu_solve, cho_factor, cho_solve
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import oracle as O
from cheq_b200 import workloads

mol = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
Z, xyz = workloads.water_cluster(mol, workloads.SEED)
P = O.default_parameters()
opt = O.Options()
chi, hard, rad, nq = O.gather(Z, P)
n = len(Z)
t0 = time.time()
M, rhs, hidx = O.build_system(xyz, chi, hard, rad, nq, 0.0, None, opt)
print("assembled", n, time.time() - t0, flush=True)
A = M[:n, :n]
c = rhs[:n].copy()
Q = rhs[n]
hmask = np.zeros(n, bool); hmask[hidx] = True
xi = np.where(~hmask)[0]; hi = np.where(hmask)[0]
Axx = np.ascontiguousarray(A[np.ix_(xi, xi)])
Axh = np.ascontiguousarray(A[np.ix_(xi, hi)])
Ahh = np.ascontiguousarray(A[np.ix_(hi, hi)])
del M, A
t0 = time.time()
lux = lu_factor(Axx)
AxxiAxh = lu_solve(lux, Axh)
S0 = Ahh - Axh.T @ AxxiAxh
S0 = 0.5 * (S0 + S0.T)
ax1 = lu_solve(lux, np.ones(len(xi))); axc = lu_solve(lux, c[xi])
t1 = np.ones(len(hi)) - Axh.T @ ax1
tc = c[hi] - Axh.T @ axc
g = ax1.sum(); h = axc.sum()
print("schur", time.time() - t0, flush=True)
nh = len(hi)
Jh = hard[hi]
d0 = np.diag(S0).copy()

def full_from_h(xh, mu):
    xx = axc - mu * ax1 - AxxiAxh @ xh
    x = np.empty(n); x[xi] = xx; x[hi] = xh
    return x

class Factor:
    def __init__(self, D):
        S = S0.copy(); S[np.diag_indices(nh)] = d0 + D
        self.lu = lu_factor(S, overwrite_a=True)
        self.D = D.copy()
        self.w = lu_solve(self.lu, t1)
        self.den = t1 @ self.w + g
    def apply(self, r, rho):           # K_s^-1 [r; rho]
        u0 = lu_solve(self.lu, r)
        nu = (t1 @ u0 - rho) / self.den
        return u0 - nu * self.w, nu

def Kmul(D, u, nu):
    return S0 @ u + D * u + nu * t1, t1 @ u - g * nu

def gmres_ir(F, D, b, beta, x0, mu0, tol=2e-15, mmax=40):
    """returns x, mu, total preconditioned matvecs"""
    x, mu = x0.copy(), mu0
    bn = np.sqrt(b @ b + beta * beta)
    steps = 0
    hist = []
    for outer in range(6):
        ru, rn = Kmul(D, x, mu)
        ru = b - ru; rn = beta - rn
        res = np.sqrt(ru @ ru + rn * rn)
        hist.append(res / bn)
        if res <= tol * bn * 50 or (outer > 0 and hist[-1] > 0.5 * hist[-2]):
            break
        # GMRES on left-preconditioned system P^-1 K d = P^-1 r
        zu, zn = F.apply(ru, rn)
        z = np.concatenate([zu, [zn]])
        beta0 = np.linalg.norm(z)
        V = [z / beta0]
        H = np.zeros((mmax + 1, mmax))
        gvec = np.zeros(mmax + 1); gvec[0] = beta0
        cs, sn = [], []
        k = 0
        for j in range(mmax):
            wu, wn = Kmul(D, V[j][:nh], V[j][nh])
            wu, wn = F.apply(wu, wn)
            w = np.concatenate([wu, [wn]])
            steps += 1
            for _ in range(2):
                for i in range(j + 1):
                    hh = V[i] @ w
                    H[i, j] += hh
                    w -= hh * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            V.append(w / H[j + 1, j] if H[j + 1, j] > 0 else w)
            for i in range(j):
                a = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = a
            rr = np.hypot(H[j, j], H[j + 1, j])
            cs.append(H[j, j] / rr); sn.append(H[j + 1, j] / rr)
            H[j, j] = rr; H[j + 1, j] = 0
            gvec[j + 1] = -sn[j] * gvec[j]; gvec[j] = cs[j] * gvec[j]
            k = j + 1
            if abs(gvec[j + 1]) <= 1e-3 * tol * beta0 / max(hist[-1], 1e-300) * 1 or abs(gvec[j + 1]) < 1e-16 * beta0 or abs(gvec[j+1]) <= tol*bn*0.0:
                break
            # stop the cycle when the preconditioned residual dropped by 1e-14 relative to the first cycle's
            if abs(gvec[j + 1]) <= 1e-15 * beta0 / max(hist[-1] / hist[0], 1e-30) * 1e0:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), gvec[:k])
        d = sum(y[i] * V[i] for i in range(k))
        x += d[:nh]; mu += d[nh]
    return x, mu, steps, hist

# --- reference SCF trajectory with exact solves; at each iteration try stale factors from every earlier iteration
charges = np.zeros(n)
damping = 0.4; prev = np.finfo(float).max
factors = {}
keep = {1, 2, 3, 4, 5, 6, 8}
xh_prev = np.zeros(nh); mu_prev = 0.0
report = []
for it in range(1, 60):
    qcl = np.clip(charges[hi], -0.95, 0.95)
    D = Jh * qcl * O.H_CHARGE_DEPENDENCE_FACTOR
    t0 = time.time()
    Fk = Factor(D)
    xh, mu = Fk.apply(tc, Q - h)
    # one refinement
    xh, mu, _, hist = gmres_ir(Fk, D, tc, Q - h, xh, mu)
    tf = time.time() - t0
    row = {"it": it, "damping": damping, "res_direct": hist[0], "t_factor": tf}
    for j, Fj in factors.items():
        dmax = float(np.max(np.abs(D - Fj.D)))
        xs, mus, steps, hist2 = gmres_ir(Fj, D, tc, Q - h, xh_prev, mu_prev)
        err = float(max(np.max(np.abs(xs - xh)), abs(mus - mu)))
        row[f"stale{j}"] = {"dmax": dmax, "steps": steps, "err": err, "res": hist2}
    x = full_from_h(xh, mu)
    delta = float(np.max(np.abs(x - charges)))
    charges = (1 - damping) * charges + damping * x
    row["delta"] = delta
    print(json.dumps(row), flush=True)
    if it in keep:
        factors[it] = Fk
    xh_prev, mu_prev = xh, mu
    if delta < opt.tolerance:
        break
    if delta > prev: damping *= 0.5
    elif delta < 0.9 * prev: damping = min(damping * 1.1, 1.0)
    damping = max(damping, 0.001)
    prev = delta
print("iterations", it, "mu", mu)
