"""Seeded synthetic inputs of the benchmark configurations (BASELINE.json configs).

Pure host-side generators (numpy); the same arrays feed the CPU oracle and the GPU path.
"""
from __future__ import annotations

import math

import numpy as np

SEED = 20261019


def readme_water():
    """W3: the README geometry (/root/reference/README.md:77-87)."""
    b, a = 0.9575, math.radians(104.45)
    return [8, 1, 1], np.array([[0.0, 0.0, 0.0], [b, 0.0, 0.0], [b * math.cos(a), b * math.sin(a), 0.0]])


def _random_rotations(rng, count):
    """Haar-random rotation matrices from unit quaternions."""
    q = rng.normal(size=(count, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((count, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def water_cluster(n_molecules: int, seed: int = SEED, spacing: float = 3.1, jitter: float = 0.15):
    """Open (non-periodic) cluster of rigid waters on a jittered cubic lattice, random orientations.

    Returns (Z (3M,), xyz (3M, 3)); atom order O, H, H per molecule."""
    rng = np.random.default_rng(seed)
    side = math.ceil(n_molecules ** (1.0 / 3.0) - 1e-9)
    grid = np.array([(i, j, k) for i in range(side) for j in range(side) for k in range(side)], dtype=float)
    centers = grid[:n_molecules] * spacing + rng.uniform(-jitter, jitter, size=(n_molecules, 3))
    r_oh, ang = 0.9572, math.radians(104.52)
    local = np.array([[0.0, 0.0, 0.0],
                      [r_oh * math.sin(ang / 2), 0.0, r_oh * math.cos(ang / 2)],
                      [-r_oh * math.sin(ang / 2), 0.0, r_oh * math.cos(ang / 2)]])
    R = _random_rotations(rng, n_molecules)
    xyz = centers[:, None, :] + np.einsum("mij,aj->mai", R, local)
    Z = np.tile(np.array([8, 1, 1]), n_molecules)
    return Z.tolist(), xyz.reshape(-1, 3)


def water_box_3k(seed: int = SEED):
    """B3k: 1,000 waters + uniform field + 10,000 embedding point charges on a shell outside the box."""
    Z, xyz = water_cluster(1000, seed)
    rng = np.random.default_rng(seed + 1)
    m = 10000
    center = xyz.mean(axis=0)
    half = 0.5 * (xyz.max(axis=0) - xyz.min(axis=0)).max()
    d = rng.normal(size=(m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    radius = half * math.sqrt(3.0) + rng.uniform(4.0, 12.0, size=m)
    pc_xyz = center + d * radius[:, None]
    pc_Z = rng.choice([1, 6, 7, 8], size=m)
    pc_q = rng.uniform(-0.8, 0.8, size=m)
    pc_q -= pc_q.mean()
    field = [0.05, 0.0, 0.1]
    return Z, xyz, {"Z": pc_Z.tolist(), "xyz": pc_xyz, "q": pc_q, "field": field}


def organic_chain(n_atoms: int = 1000, seed: int = SEED):
    """MD base structure: a self-avoiding heavy-atom chain (C/N/O) decorated with hydrogens.

    Composition approx C 32 %, H 50 %, N 6 %, O 12 %; bonded distances 1.0-1.5 Angstrom, non-bonded >= 1.6."""
    rng = np.random.default_rng(seed + 2)
    n_heavy = n_atoms // 2
    n_h = n_atoms - n_heavy
    heavy_Z = rng.choice([6, 7, 8], size=n_heavy, p=[0.64, 0.12, 0.24])
    pos = [np.zeros(3)]
    cell = {}

    def key(p):
        return tuple(np.floor(p / 1.6).astype(int))

    def clash(p, min_d, skip=()):
        k = key(p)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for idx in cell.get((k[0] + dx, k[1] + dy, k[2] + dz), ()):
                        if idx in skip:
                            continue
                        if np.linalg.norm(all_pos[idx] - p) < min_d:
                            return True
        return False

    all_pos = [pos[0]]
    cell.setdefault(key(pos[0]), []).append(0)
    direction = np.array([1.0, 0.0, 0.0])
    while len(all_pos) < n_heavy:
        for _ in range(200):
            d = direction + 0.9 * rng.normal(size=3)
            d /= np.linalg.norm(d)
            p = all_pos[-1] + d * rng.uniform(1.38, 1.5)
            if not clash(p, 1.6, skip=(len(all_pos) - 1,)):
                direction = d
                break
        else:
            direction = rng.normal(size=3)
            direction /= np.linalg.norm(direction)
            continue
        cell.setdefault(key(p), []).append(len(all_pos))
        all_pos.append(p)
    Z = list(heavy_Z)
    placed = 0
    host = 0
    attempts = 0
    while placed < n_h:
        parent = host % n_heavy
        host += 1
        attempts += 1
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        p = all_pos[parent] + d * rng.uniform(1.0, 1.1)
        if clash(p, 1.6 if attempts < 50 * n_h else 1.2, skip=(parent,)):
            continue
        cell.setdefault(key(p), []).append(len(all_pos))
        all_pos.append(p)
        Z.append(1)
        placed += 1
    return [int(z) for z in Z], np.array(all_pos)


def md_frames(base_xyz, n_frames: int, seed: int = SEED, sigma: float = 0.05):
    """Frame f = base + N(0, sigma) per coordinate, seeded seed + f."""
    out = np.empty((n_frames,) + base_xyz.shape)
    for f in range(n_frames):
        out[f] = base_xyz + np.random.default_rng(seed + f).normal(scale=sigma, size=base_xyz.shape)
    return out


def md_frames_subset(base_xyz, indices, seed: int = SEED, sigma: float = 0.05):
    """The frames `indices` of md_frames(base_xyz, ...) without generating the others (one rank's shard)."""
    out = np.empty((len(indices),) + base_xyz.shape)
    for k, f in enumerate(indices):
        out[k] = base_xyz + np.random.default_rng(seed + int(f)).normal(scale=sigma, size=base_xyz.shape)
    return out


def water_20k(seed: int = SEED):
    """S20k: 6,667 waters = 20,001 atoms (the headline configuration)."""
    return water_cluster(6667, seed)


def solvated_100k(n_solute: int = 10000, n_waters: int = 30000, seed: int = SEED, spacing: float = 3.1,
                  jitter: float = 0.15, clearance: float = 2.6):
    """S100k (BASELINE config 5): an `n_solute`-atom organic chain (C/H/N/O) solvated by the `n_waters` lattice
    waters closest to it (oxygen at least `clearance` Angstrom from every solute atom) -> 100,000 atoms.
    Returns (Z, xyz); solute atoms first, then O, H, H per water."""
    from scipy.spatial import cKDTree
    Zs, xs = organic_chain(n_solute, seed)
    rng = np.random.default_rng(seed + 7)
    lo, hi = xs.min(axis=0) - 12.0, xs.max(axis=0) + 12.0
    axes = [np.arange(lo[d], hi[d], spacing) for d in range(3)]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    dist, _ = cKDTree(xs).query(grid)
    ok = np.where(dist >= clearance + jitter * math.sqrt(3.0))[0]
    if len(ok) < n_waters:
        raise ValueError("water lattice too small")
    keep = ok[np.argsort(dist[ok], kind="stable")[:n_waters]]
    keep.sort()
    centers = grid[keep] + rng.uniform(-jitter, jitter, size=(n_waters, 3))
    r_oh, ang = 0.9572, math.radians(104.52)
    local = np.array([[0.0, 0.0, 0.0],
                      [r_oh * math.sin(ang / 2), 0.0, r_oh * math.cos(ang / 2)],
                      [-r_oh * math.sin(ang / 2), 0.0, r_oh * math.cos(ang / 2)]])
    R = _random_rotations(rng, n_waters)
    xw = (centers[:, None, :] + np.einsum("mij,aj->mai", R, local)).reshape(-1, 3)
    Z = list(Zs) + [8, 1, 1] * n_waters
    return Z, np.vstack([xs, xw])


def gather_parameters(Z, parameters):
    """Z -> (chi, hardness, radius, n) arrays using a cheq_b200.Parameters (host-side element lookup)."""
    el = parameters.elements
    chi = np.array([el[int(z)].electronegativity for z in Z])
    hard = np.array([el[int(z)].hardness for z in Z])
    rad = np.array([el[int(z)].radius for z in Z])
    nq = np.array([el[int(z)].principal_quantum_number for z in Z], dtype=np.uint8)
    return chi, hard, rad, nq
