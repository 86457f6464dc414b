"""CPU tests of the multi-GPU factorisation's host-side logic: the panel/owner map the C library computes
(cheq_dist_plan), the communicator-id rendezvous, and -- over a world-size-2 gloo group -- a NumPy model of the
distributed schedule (tests/dist_model.py) against the bordered LU of the reference
(/root/reference/src/solver/implementation.rs:557-603, restated in oracle/)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def test_plan_partition_and_balance():
    from cheq_b200 import distributed
    for xt, ht in ((53, 105), (1, 1), (0, 7), (9, 0), (3, 4)):
        for tpp in (1, 2, 4, 5):
            for world in (1, 2, 4, 8):
                owner, panel = distributed.plan(xt, ht, tpp, world)
                assert owner.min() >= 0 and owner.max() < world and panel.min() == 0
                assert np.all(np.diff(panel) >= 0) and np.all(np.diff(panel) <= 1)      # panels are contiguous
                for p in range(panel.max() + 1):
                    t = np.where(panel == p)[0]
                    assert 1 <= len(t) <= tpp and len(set(owner[t])) == 1                # one owner per panel
                    assert (t[-1] < xt) or (t[0] >= xt)                                  # never straddles X | H
                    assert owner[t[0]] == p % world                                      # block-cyclic
                counts = np.bincount(owner, minlength=world)
                if xt + ht >= 2 * world * tpp:
                    assert counts.max() - counts.min() <= 2 * tpp


def test_plan_rejects_bad_arguments():
    from cheq_b200 import _ffi
    lib = _ffi.lib()
    assert lib.cheq_dist_plan(4, 4, 0, 2, None, None) == _ffi.INVALID_ARGUMENT
    assert lib.cheq_dist_plan(4, 4, 2, 0, None, None) == _ffi.INVALID_ARGUMENT
    assert lib.cheq_dist_plan(-1, 4, 2, 2, None, None) == _ffi.INVALID_ARGUMENT


def _system(n_mol, seed=3):
    from cheq_b200 import workloads
    from oracle import oracle as O
    Z, xyz = workloads.water_cluster(n_mol, seed)
    P = O.default_parameters()
    chi, hard, rad, nq = O.gather(Z, P)
    M, rhs, hidx = O.build_system(xyz, chi, hard, rad, nq, 0.0, None, O.Options())
    n = len(Z)
    return Z, xyz, np.array(M[:n, :n]), np.asarray(chi), np.asarray(hard), np.asarray(nq)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from cheq_b200 import distributed
    from dist_model import dist_solve_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rendezvous of the communicator id: every rank ends up with rank 0's 128 bytes
    raw = distributed.exchange_unique_id(lambda: bytes(range(128)))
    assert raw == bytes(range(128))

    def bcast(a, root):
        t = torch.from_numpy(a)
        dist.broadcast(t, src=root)

    Z, xyz, A, chi, hard, nq = _system(14)
    tile = 8
    nx, nh = int((nq != 1).sum()), int((nq == 1).sum())
    up = lambda v: (v + tile - 1) // tile
    for tpp in (1, 2):
        owner, panel = distributed.plan(up(nx), up(nh), tpp, world)
        q, mu, its = dist_solve_model(A, chi, hard, nq, 0.0, rank, world, bcast, owner, panel, tile)
        np.save(Path(out_dir) / f"q{tpp}_{rank}.npy", np.concatenate([q, [mu, its]]))
    dist.destroy_process_group()


def test_distributed_schedule_model_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    port = 29500 + (os.getpid() + 977) % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    Z, xyz, *_ = _system(14)
    ref = O.solve(Z, xyz, 0.0)
    for tpp in (1, 2):
        a, b = np.load(tmp_path / f"q{tpp}_0.npy"), np.load(tmp_path / f"q{tpp}_1.npy")
        assert np.array_equal(a, b)                                   # replicated solves agree bit for bit
        assert int(a[-1]) == ref.iterations
        assert np.max(np.abs(a[:-2] - ref.charges)) < 1e-10 and abs(a[-2] - ref.equilibrated_potential) < 1e-10


def test_distributed_schedule_model_single_rank():
    from cheq_b200 import distributed
    from dist_model import dist_solve_model
    from oracle import oracle as O
    Z, xyz, A, chi, hard, nq = _system(9, seed=5)
    tile = 4
    up = lambda v: (v + tile - 1) // tile
    owner, panel = distributed.plan(up(int((nq != 1).sum())), up(int((nq == 1).sum())), 3, 1)
    q, mu, its = dist_solve_model(A, chi, hard, nq, 0.0, 0, 1, lambda a, r: None, owner, panel, tile)
    ref = O.solve(Z, xyz, 0.0)
    assert its == ref.iterations and np.max(np.abs(q - ref.charges)) < 1e-10
