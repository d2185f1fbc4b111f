#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid) into a markdown table.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md

The per-launch times of such a pass are cold-cache and serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki]
        for pre in ("void ", "mmd::"):
            if name.startswith(pre):
                name = name[len(pre):]
        name = name.split("(")[0]
        agg[(name[:70], r[gi])].append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | grid | launches | total us | avg us | share |")
    print("|---|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| `%s` | %s | %d | %.1f | %.1f | %.3f |" % (k[0], k[1], len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, sum(v) / tot))
    print()
    print("total device time in the list: %.1f us over %d launches" % (tot / 1e3, sum(len(v) for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
