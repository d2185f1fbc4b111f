#!/usr/bin/env python
"""Executed-instruction histogram per SASS opcode (+ stall-sample totals) of one kernel in an .ncu-rep.

    python tools/ncu_opcodes.py prof.ncu-rep <kernel regex> <launch index> [tiles]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, regex, idx = sys.argv[1], sys.argv[2], sys.argv[3]
tiles = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:%s" % (regex, idx)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
print(rows[hi - 1][1] if hi else "")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops, samples, st = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = ".".join((m.group(2) if m else src).split(".")[:3])
    n = int(r[ix["Instructions Executed"]])
    ops[op] += n
    tot += n
    samples[op] += int(r[ix["# Samples"]])
    for h in hdr:
        if h.startswith("stall_") and "Not" not in h:
            st[h] += int(r[ix[h]])
print("total warp instructions %d (%.1f per tile)" % (tot, tot / tiles))
for op, n in ops.most_common(45):
    print("%-28s %10d %8.1f/tile %5.1f%%  samples %d" % (op, n, n / tiles, 100.0 * n / tot, samples[op]))
print(st.most_common())
