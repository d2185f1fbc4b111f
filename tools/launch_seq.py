#!/usr/bin/env python
"""Print the launch sequence (kernel, grid, us) of an `ncu --metrics gpu__time_duration.sum --csv` log, optionally only
the last N launches:  python tools/launch_seq.py launches.csv [N] [name filter regex]"""
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
    rows.append((re.sub(r"^void |mmd::", "", r["Kernel Name"])[:60], r["Grid Size"], us))
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows)
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
for name, grid, us in rows[-n:]:
    if pat is None or pat.search(name):
        print("%-62s %-16s %8.1f" % (name, grid, us))
