"""Multi-GPU parity worker: run with `python -m torch.distributed.run --nproc-per-node N tests/dist_worker.py`.

Every rank joins one engine group (cheq_ctx_dist_init) and calls the collective single-frame solve with identical
inputs.  Rank 0 checks the result against the CPU oracle (small cases) and against a non-distributed context on
its own GPU (larger cases); every rank checks that it got the same charges as rank 0.  Exit code 0 = pass."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist

    import cheq_b200
    from cheq_b200 import QEqSolver, SolverOptions, BasisType, _ffi, distributed, get_default_parameters, workloads

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    ctx = cheq_b200.Context(local)
    distributed.join(ctx)
    r_, w_ = C.c_int32(), C.c_int32()
    ctx._lib.cheq_ctx_dist_info(ctx.handle, C.byref(r_), C.byref(w_))
    assert (r_.value, w_.value) == (rank, world)
    local_ctx = cheq_b200.Context(local) if rank == 0 else None   # plain single-GPU engine for comparisons
    if local_ctx is not None:
        local_ctx._lib.cheq_set_factor_mode(local_ctx.handle, 2, 0)   # recursive schedule, whatever the environment says
    P = get_default_parameters()
    big = int(os.environ.get("CHEQ_DIST_TEST_BIG", "2200"))

    def check_same_everywhere(q, mu, its):
        t = torch.tensor(np.concatenate([q, [mu, float(its)]]), device=device)
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), "ranks disagree"

    def run_case(name, Z, xyz, opts, ext=None, oracle_check=True, tpp=4, expect_fallback=None):
        ctx._lib.cheq_set_factor_mode(ctx.handle, 1 if world == 1 else 0, tpp)
        chi, hard, rad, nq = workloads.gather_parameters(Z, P)
        s = QEqSolver(P, opts, ctx)
        res = s.solve_arrays(np.ascontiguousarray(xyz), chi, hard, rad, nq, 0.0, ext)
        q = np.array(res.charges)
        check_same_everywhere(q, res.equilibrated_potential, res.iterations)
        st = ctx.stats()
        if expect_fallback is not None:
            assert st["lu_fallback_frames"] == expect_fallback, st["lu_fallback_frames"]
        if rank == 0:
            single = QEqSolver(P, opts, local_ctx).solve_arrays(np.ascontiguousarray(xyz), chi, hard, rad, nq, 0.0, ext)
            dq = float(np.max(np.abs(q - np.array(single.charges))))
            dmu = abs(res.equilibrated_potential - single.equilibrated_potential)
            assert res.iterations == single.iterations, (res.iterations, single.iterations)
            assert dq < 1e-9 and dmu < 1e-9, (name, dq, dmu)
            msg = f"{name}: n={len(Z)} world={world} tpp={tpp} its={res.iterations} vs single-GPU dq={dq:.1e} dmu={dmu:.1e}"
            if oracle_check:
                from oracle import oracle as O
                oo = O.Options(tolerance=opts.tolerance, max_iterations=opts.max_iterations,
                               lambda_scale=opts.lambda_scale, hydrogen_scf=opts.hydrogen_scf,
                               basis=opts.basis_type.value, damping=(opts.damping.kind, opts.damping.value))
                ref = O.solve(list(Z), np.asarray(xyz), 0.0, oo)
                dqo = float(np.max(np.abs(q - ref.charges)))
                dmuo = abs(res.equilibrated_potential - ref.equilibrated_potential)
                assert res.iterations == ref.iterations and dqo <= 1e-8 and dmuo <= 1e-8, (name, dqo, dmuo)
                msg += f"; vs oracle dq={dqo:.1e} dmu={dmuo:.1e}"
            print(msg, flush=True)

    # small cases against the oracle: panel boundaries inside and across the X / H blocks
    Z, xyz = workloads.water_cluster(250)                      # 750 atoms: 2 X tiles + 4 H tiles
    for tpp in (1, 2, 4):
        run_case("water750/STO/scf", Z, xyz, SolverOptions(), tpp=tpp)
    run_case("water750/GTO/scf", Z, xyz, SolverOptions(basis_type=BasisType.Gto), tpp=2)
    run_case("water750/STO/no-scf", Z, xyz, SolverOptions(hydrogen_scf=False), tpp=1)
    Zw, xw = workloads.readme_water()
    run_case("water3", Zw, xw, SolverOptions())                # fewer panels than ranks
    # field + point charges (BASELINE config 3), compared with the single-GPU engine
    Z3, x3, ext = workloads.water_box_3k()
    _, _, prad, pnq = workloads.gather_parameters(ext["Z"], P)
    e = _ffi.External()
    pxyz, pq = np.ascontiguousarray(ext["xyz"]), np.ascontiguousarray(ext["q"])
    e.num_point_charges = len(pq)
    e.xyz, e.charge = pxyz.ctypes.data_as(C.c_void_p), pq.ctypes.data_as(C.c_void_p)
    e.radius, e.n = prad.ctypes.data_as(C.c_void_p), pnq.ctypes.data_as(C.c_void_p)
    e.field[0], e.field[1], e.field[2] = ext["field"]
    run_case("B3k/field+charges", Z3, x3, SolverOptions(), ext=e, oracle_check=False)
    # the reference's indefinite C2H4 fixture and a vanishing pivot (pivoted-LU path, replicated on every rank)
    import json
    fx = json.loads((ROOT / "tests" / "golden" / "rappe_goddard.json").read_text())
    c = [c for g in fx["groups"] for c in g["cases"] if c["name"] == "C2H4"][0]
    run_case("C2H4/indefinite", c["Z"], np.array(c["xyz"]), SolverOptions(), expect_fallback=0)
    # batches on a JOINED context stay local whatever the shard size (a one-frame shard must not enter
    # the collective factorisation): world + 1 frames -> rank 0 gets two frames, every other rank exactly one
    from cheq_b200.parallel import solve_batch_sharded
    Zm, base = workloads.organic_chain(200)
    frames = workloads.md_frames(base, world + 1)
    sb = QEqSolver(P, SolverOptions(), ctx)
    qb, mub, itb, stb = solve_batch_sharded(lambda fr: sb.solve_batch(Zm, fr, 0.0), frames)
    assert np.all(stb == 0)
    if rank == 0:
        from oracle import oracle as O
        for f in range(len(frames)):
            ref = O.solve(list(Zm), frames[f], 0.0)
            assert itb[f] == ref.iterations and np.max(np.abs(qb[f] - ref.charges)) <= 1e-8, f
        print(f"batch of {world + 1} frames on the joined contexts: per-frame oracle parity ok", flush=True)
    # ... and the next collective solve still works afterwards
    run_case("water750/after-batch", Z, xyz, SolverOptions(), tpp=2)
    # a larger system against the single-GPU engine
    Zb, xb = workloads.water_cluster(big)
    run_case(f"water{3 * big}", Zb, xb, SolverOptions(), oracle_check=False)
    # stale-factor SCF iterations on the group: S0 by tile columns, R by row chunks, three allreduces per GMRES step
    ctx.set_krylov(True, 128, 4.5, 5)
    run_case("water750/STO/scf/stale-factor", Z, xyz, SolverOptions(), tpp=2)
    assert ctx.stats()["krylov_iterations"] > 0
    run_case("water750/GTO/scf/stale-factor", Z, xyz, SolverOptions(basis_type=BasisType.Gto), tpp=1)
    run_case(f"water{3 * big}/stale-factor", Zb, xb, SolverOptions(), oracle_check=False)
    assert ctx.stats()["krylov_iterations"] > 0
    ctx.set_krylov(True, 4096, -1.0, 5)
    # matrix-free path on the group: rows of the operator sharded, one all-gather per application
    chi, hard, rad, nq = workloads.gather_parameters(Z, P)
    s_it = QEqSolver(P, SolverOptions(), ctx)
    res = s_it.solve_arrays(np.ascontiguousarray(xyz), chi, hard, rad, nq, 0.0, None, iterative=True)
    check_same_everywhere(np.array(res.charges), res.equilibrated_potential, res.iterations)
    if rank == 0:
        from oracle import oracle as O
        ref = O.solve(list(Z), np.asarray(xyz), 0.0)
        dqo = float(np.max(np.abs(np.array(res.charges) - ref.charges)))
        assert res.iterations == ref.iterations and dqo <= 1e-8, (res.iterations, ref.iterations, dqo)
        print(f"water750/matrix-free on {world} rank(s): its={res.iterations} vs oracle dq={dqo:.1e}, "
              f"{ctx.stats()['matfree_matvecs']} operator applications", flush=True)
    dist.barrier()
    if rank == 0:
        print("DIST PARITY PASS", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
