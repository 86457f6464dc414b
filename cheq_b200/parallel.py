"""Multi-GPU sharding of batched frames (frames are independent units, one process per GPU,
no data-path collective).  Rank r solves frames r, r + world, r + 2 world, ...; results are gathered with one
all_gather of padded blocks (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_indices(n_frames: int, rank: int, world: int) -> np.ndarray:
    """Frame indices owned by `rank` (round-robin, so ranks stay balanced for any n_frames)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    return np.arange(rank, n_frames, world, dtype=np.int64)


def solve_batch_sharded(solve_fn, frames_xyz, group=None):
    """Solves a trajectory across the ranks of `group` (default process group).

    solve_fn(local_frames (f, N, 3)) -> (charges (f, N), potentials (f,), iterations (f,), status (f,)).
    Returns the full-trajectory arrays on every rank.  Without an initialised process group it runs locally."""
    import torch
    import torch.distributed as dist

    frames_xyz = np.ascontiguousarray(frames_xyz, dtype=np.float64)
    F, n = frames_xyz.shape[0], frames_xyz.shape[1]
    if not (dist.is_available() and dist.is_initialized()):
        return solve_fn(frames_xyz)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_indices(F, rank, world)
    q, mu, its, status = solve_fn(frames_xyz[mine]) if len(mine) else (np.zeros((0, n)), np.zeros(0), np.zeros(0), np.zeros(0))
    per = (F + world - 1) // world
    block = np.zeros((per, n + 3))
    block[:len(mine), :n] = q
    block[:len(mine), n] = mu
    block[:len(mine), n + 1] = its
    block[:len(mine), n + 2] = status
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    send = torch.from_numpy(block).to(device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    charges = np.empty((F, n))
    pot = np.empty(F)
    iters = np.empty(F, dtype=np.uint32)
    stat = np.empty(F, dtype=np.int32)
    for r in range(world):
        idx = shard_indices(F, r, world)
        b = recv[r].cpu().numpy()[:len(idx)]
        charges[idx] = b[:, :n]
        pot[idx] = b[:, n]
        iters[idx] = b[:, n + 1].astype(np.uint32)
        stat[idx] = b[:, n + 2].astype(np.int32)
    return charges, pot, iters, stat
