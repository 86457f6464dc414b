"""One dense QEq system across the GPUs of a node (SURVEY.md section 8(e), row "direct factorisation of one big
system").  The reference has no counterpart (single-node rayon, /root/reference/src/solver/implementation.rs:394,
491); the host side here only does the rendezvous: rank 0 asks the engine for a communicator id, the 128 bytes
travel through the host process group (NCCL on GPUs, gloo in the CPU tests), and every rank joins with its own
context.  All numerical work -- owner-computes assembly and trailing updates, panel broadcasts over NVLink,
replicated triangular solves -- is behind the C ABI (cheq_ctx_dist_init, include/cheq_b200.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi

ID_BYTES = 128


def exchange_unique_id(make_id, group=None) -> bytes:
    """Rank 0 creates the id with `make_id()` (-> 128 bytes); every rank returns the same bytes."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if rank == 0:
        raw = bytes(make_id())
        if len(raw) != ID_BYTES:
            raise ValueError("communicator id must be 128 bytes")
        t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(ID_BYTES, dtype=torch.uint8, device=device)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def engine_unique_id() -> bytes:
    buf = (C.c_uint8 * ID_BYTES)()
    rc = _ffi.lib().cheq_dist_unique_id(buf)
    if rc != _ffi.OK:
        raise RuntimeError(f"cheq_dist_unique_id failed with status {rc} (is NCCL loadable? set CHEQ_NCCL_LIB)")
    return bytes(buf)


def join(ctx, group=None) -> None:
    """Makes `ctx` (one Context per process, on this process's GPU) part of the process group: afterwards
    single-frame solves on it are collective over all ranks of the group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    rank = dist.get_rank(group)
    raw = exchange_unique_id(engine_unique_id, group)
    buf = (C.c_uint8 * ID_BYTES).from_buffer_copy(raw)
    rc = ctx._lib.cheq_ctx_dist_init(ctx.handle, rank, world, buf)
    if rc != _ffi.OK:
        raise RuntimeError(f"cheq_ctx_dist_init failed with status {rc}: {ctx.last_error()}")


def plan(x_tiles: int, h_tiles: int, tiles_per_panel: int, world: int):
    """(owner, panel) of every 128-column tile, as the engine computes them (host-only call)."""
    n = x_tiles + h_tiles
    owner = np.full(n, -1, dtype=np.int32)
    panel = np.full(n, -1, dtype=np.int32)
    rc = _ffi.lib().cheq_dist_plan(x_tiles, h_tiles, tiles_per_panel, world, owner.ctypes.data_as(C.c_void_p),
                                   panel.ctypes.data_as(C.c_void_p))
    if rc != _ffi.OK:
        raise ValueError(f"cheq_dist_plan: status {rc}")
    return owner, panel
