"""World-size-2 gloo test (CPU) of the multi-GPU frame sharding: every frame solved exactly once, results gathered
in trajectory order on every rank.  The per-rank solve is a stand-in (host logic only); the GPU solve itself is
covered by the -m gpu tests."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _fake_solve(frames):
    f, n = frames.shape[0], frames.shape[1]
    q = frames.sum(axis=2)                          # any deterministic per-frame function
    mu = frames.reshape(f, -1).mean(axis=1)
    its = 9 + (np.abs(frames[:, 0, 0]) * 10).astype(int) % 3   # depends on the frame, not on its local index
    status = np.zeros(f)
    return q, mu, its, status


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from cheq_b200.parallel import shard_indices, solve_batch_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    frames = rng.normal(size=(n_frames, 7, 3))
    solved = []

    def solve(local):
        solved.append(len(local))
        return _fake_solve(local)
    q, mu, its, status = solve_batch_sharded(solve, frames)
    ref = _fake_solve(frames)
    assert np.array_equal(q, ref[0]) and np.array_equal(mu, ref[1])
    assert np.array_equal(its, ref[2].astype(np.uint32)) and np.all(status == 0)
    mine = len(shard_indices(n_frames, rank, world))
    assert solved == ([mine] if mine else [])    # a rank without frames does not call the solver
    np.save(Path(out_dir) / f"rank{rank}.npy", q)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8, 1])
def test_frame_sharding_world2_gloo(tmp_path, n_frames):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + n_frames) % 2000
    mp.spawn(_worker, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b) and a.shape == (n_frames, 7)


def test_shard_indices_partition():
    from cheq_b200.parallel import shard_indices
    for F in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            got = np.sort(np.concatenate([shard_indices(F, r, world) for r in range(world)]))
            assert np.array_equal(got, np.arange(F))
            sizes = [len(shard_indices(F, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
