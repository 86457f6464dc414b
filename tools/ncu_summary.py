#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("cheq::<unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':42s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:42]:42s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%")
    print(f"{'total':42s} {sum(v[0] for v in agg.values()):8d} {tot:12.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
