#!/usr/bin/env python3
"""CPU experiment: GMRES iteration counts for the bordered QEq system with a block-Jacobi preconditioner over
spatially compact (Morton-ordered) atom blocks, with and without a coarse space of block-constant vectors.
Calibrates the matrix-free path."""
import sys, time
from pathlib import Path
import numpy as np
from scipy.linalg import lu_factor, lu_solve
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import oracle as O
from cheq_b200 import workloads

mol = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
Z, xyz = workloads.water_cluster(mol, workloads.SEED)
P = O.default_parameters(); opt = O.Options()
chi, hard, rad, nq = O.gather(Z, P)
n = len(Z)
M, rhs, hidx = O.build_system(xyz, chi, hard, rad, nq, 0.0, None, opt)
A = M[:n, :n].copy(); c = rhs[:n].copy()
del M
# Morton order over all atoms
lo = xyz.min(0); ext = (xyz.max(0) - lo).max()
qz = np.minimum(1023, ((xyz - lo) / ext * 1024).astype(np.int64))
def spread(v):
    v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249
    return v
code = spread(qz[:, 0]) | (spread(qz[:, 1]) << 1) | (spread(qz[:, 2]) << 2)
perm = np.argsort(code, kind="stable")
A = A[np.ix_(perm, perm)]; c = c[perm]
ones = np.ones(n)
xref = None

def gmres(apply_A, apply_P, b, tol=1e-12, m=200):
    x = np.zeros_like(b); total = 0
    for cycle in range(10):
        r = apply_P(b - apply_A(x)); beta = np.linalg.norm(r)
        if cycle == 0: beta0 = np.linalg.norm(apply_P(b))
        if beta <= tol * beta0: return x, total
        V = [r / beta]; H = np.zeros((m + 1, m)); g = np.zeros(m + 1); g[0] = beta; cs = []; sn = []
        for j in range(m):
            w = apply_P(apply_A(V[j])); total += 1
            for _ in range(2):
                for i in range(j + 1):
                    h = V[i] @ w; H[i, j] += h; w -= h * V[i]
            hn = np.linalg.norm(w); V.append(w / hn)
            for i in range(j):
                a = cs[i] * H[i, j] + sn[i] * H[i + 1, j]; H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]; H[i, j] = a
            rr = np.hypot(H[j, j], hn); cs.append(H[j, j] / rr); sn.append(hn / rr); H[j, j] = rr
            g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j]
            if abs(g[j + 1]) <= tol * beta0: break
        k = j + 1
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]); x = x + sum(y[i] * V[i] for i in range(k))
    return x, total

def bordered_ops(A):
    def apply_A(v):
        return np.concatenate([A @ v[:n] + v[n], [v[:n].sum()]])
    return apply_A

apply_A = bordered_ops(A)
b = np.concatenate([c, [0.0]])
for B in [int(s) for s in (sys.argv[2:] or ["512", "2048"])]:
    blocks = [(s, min(n, s + B)) for s in range(0, n, B)]
    t0 = time.time()
    lus = [lu_factor(A[s:e, s:e]) for s, e in blocks]
    def Pinv(r):
        return np.concatenate([lu_solve(lu, r[s:e]) for lu, (s, e) in zip(lus, blocks)])
    w1 = Pinv(ones); den = ones @ w1
    def apply_P(r):     # bordered block-Jacobi
        u0 = Pinv(r[:n]); nu = (ones @ u0 - r[n]) / den
        return np.concatenate([u0 - nu * w1, [nu]])
    x1, it1 = gmres(apply_A, apply_P, b)
    # two-level: coarse space of block-constant vectors (+ border handled by the bordered elimination)
    Zc = np.zeros((n, len(blocks)))
    for k, (s, e) in enumerate(blocks): Zc[s:e, k] = 1.0
    AZ = A @ Zc; E = Zc.T @ AZ; Ei = np.linalg.inv(E)
    def Pinv2(r):       # additive: block-Jacobi + coarse correction
        return Pinv(r) + Zc @ (Ei @ (Zc.T @ r))
    w2 = Pinv2(ones); den2 = ones @ w2
    def apply_P2(r):
        u0 = Pinv2(r[:n]); nu = (ones @ u0 - r[n]) / den2
        return np.concatenate([u0 - nu * w2, [nu]])
    x2, it2 = gmres(apply_A, apply_P2, b)
    # deflated (multiplicative) variant: P = Pj (I - A Z E^-1 Z^T) + Z E^-1 Z^T
    def Pinv3(r):
        cr = Zc @ (Ei @ (Zc.T @ r))
        return Pinv(r - A @ cr) + cr
    w3 = Pinv3(ones); den3 = ones @ w3
    def apply_P3(r):
        u0 = Pinv3(r[:n]); nu = (ones @ u0 - r[n]) / den3
        return np.concatenate([u0 - nu * w3, [nu]])
    x3, it3 = gmres(apply_A, apply_P3, b)
    print(f"n={n} B={B} blocks={len(blocks)}: block-Jacobi {it1} its; +coarse additive {it2}; +coarse multiplicative {it3}; "
          f"diff {np.max(np.abs(x1-x2)):.1e} {np.max(np.abs(x1-x3)):.1e}  ({time.time()-t0:.0f}s)", flush=True)
