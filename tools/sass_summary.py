#!/usr/bin/env python3
"""SASS instruction counts per kernel of cheq_b200/lib/libcheq_b200.so (cuobjdump -sass | c++filt), the evidence for
which hardware paths each kernel uses: DMMA (FP64 tensor), LDGSTS (cp.async), and the Blackwell opcodes of the
tcgen05 slice-product kernel -- UTCIMMA (tcgen05.mma.kind::i8), UTCBAR (tcgen05.commit), LDTM (tcgen05.ld),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), UTCATOMSWS (tcgen05.alloc / dealloc).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
COLS = ["DMMA", "DFMA", "LDGSTS", "LDG", "STG", "LDS", "SHFL", "MUFU", "BAR", "UTCIMMA", "UTCBAR", "LDTM", "UBLKCP",
        "SYNCS", "UTCATOMSWS"]


def main():
    so = ROOT / "cheq_b200" / "lib" / "libcheq_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for c in COLS:
                if op == c or op.startswith(c + ".") or (c == "BAR" and op == "BAR") or (c == "LDTM" and op.startswith("LDTM")):
                    counts[name][c] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass cheq_b200/lib/libcheq_b200.so (sm_100a only), SASS instruction counts per kernel (c++filt names)")
    print("# FP64 tensor path: DMMA.8x8x4 (mma.sync.m8n8k4.f64), operands by LDGSTS (cp.async).")
    print("# tcgen05 path (oz_gemm_kernel<S>, FP64-equivalent int8 slice products): UTCIMMA = tcgen05.mma.kind::i8, UTCBAR =")
    print("# tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (global -> shared, mbarrier completion), SYNCS = mbarrier")
    print("# ops, UTCATOMSWS = tcgen05.alloc / dealloc / relinquish.")
    print(f"{'kernel':46s}" + "".join(f"{c:>8s}" for c in COLS))
    rows = []
    for mangled, pretty in zip(counts, names):
        short = pretty.replace("(anonymous namespace)::", "").replace("cheq::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", short)
        rows.append((short, counts[mangled]))
    for short, c in sorted(rows):
        print(f"{short[:46]:46s}" + "".join(f"{c[k]:8d}" for k in COLS))


if __name__ == "__main__":
    sys.exit(main())
