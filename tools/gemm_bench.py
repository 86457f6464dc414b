#!/usr/bin/env python3
"""GEMM microbenchmark on the GPU box: the engine's DMMA kernel vs cuBLAS DGEMM at the shapes the
factorisation uses (tall C, K from 128 up)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import cheq_b200


def ours(ctx, m, n, k, lower, reps=5):
    ms = C.c_double()
    rc = ctx._lib.cheq_dbg_gemm_bench(ctx.handle, m, n, k, lower, 1, reps, C.byref(ms))
    assert rc == 0, ctx.last_error()
    fl = 2.0 * m * n * k if not lower else (n * n * k + 2.0 * (m - n) * n * k)
    return ms.value, fl / (ms.value * 1e-3) / 1e12


def cublas(m, n, k, reps=5):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda")
    b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    c = torch.randn(m, n, dtype=torch.float64, device="cuda")
    torch.addmm(c, a, b, alpha=-1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.addmm(c, a, b, alpha=-1.0, out=c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * m * n * k / (ms * 1e-3) / 1e12


def main():
    ctx = cheq_b200.default_context(0)
    if sys.argv[1:] == ["tiles"]:
        for m, n, k, lower in [(13440, 128, 128, 0), (6784, 128, 128, 0), (3328, 128, 128, 0), (13440, 256, 256, 1),
                               (6784, 256, 128, 0), (6784, 384, 384, 1), (13440, 384, 384, 1), (6784, 896, 768, 1),
                               (13440, 896, 768, 1), (6784, 1664, 1664, 1)]:
            out = []
            for bit, name in ((0x80, "128"), (0x40, "64")):
                ms, _ = ours(ctx, m, n, k, lower | bit)
                fl = 2.0 * m * n * k if not lower else (n * n * k + 2.0 * (m - n) * n * k)
                out.append(f"tile{name}: {ms * 1e3:8.1f} us {fl / (ms * 1e-3) / 1e12:6.2f} TF/s")
            print(f"m={m:6d} n={n:5d} k={k:5d} lower={lower}  " + " | ".join(out), flush=True)
        return
    only = sys.argv[1:] == ["one"]
    shapes = [(8192, 8192, 8192, 0)] if only else [
        (8192, 8192, 8192, 0), (13568, 13440, 6784, 1), (13568, 6656, 6784, 1), (6784, 3328, 3456, 1),
        (13440, 128, 128, 0), (13440, 128, 256, 0), (13440, 256, 256, 1), (13440, 512, 512, 1),
        (13440, 1024, 1024, 1), (6784, 128, 128, 0), (2048, 128, 128, 0)]
    for m, n, k, lower in shapes:
        ms, tf = ours(ctx, m, n, k, lower)
        line = f"m={m:6d} n={n:6d} k={k:5d} lower={lower}  ours {ms:9.3f} ms {tf:6.2f} TF/s"
        if not only:
            cms, ctf = cublas(m, n, k)
            line += f" | cuBLAS full {cms:9.3f} ms {ctf:6.2f} TF/s"
        print(line, flush=True)


if __name__ == "__main__":
    main()
