#!/usr/bin/env python
"""Per-source-line instruction / stall-sample table for one kernel of an .ncu-rep (needs -lineinfo at compile time).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel regex> <launch index> <cubin> [top N] [mangled-name substring]

ncu's CSV source page is SASS-level only; nvdisasm -g gives the line of every SASS instruction of the same cubin.
The two listings are matched by instruction order inside the function.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def sass_rows(rep, regex, idx):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:%s" % (regex, idx)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    name = rows[hi - 1][1] if hi > 0 else ""
    return name, hdr, rows[hi + 1:]


def line_map(cubin, func_substr):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    res, cur, infunc = [], None, False
    for ln in lines:
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infunc = func_substr in m.group(1)
            continue
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            res.append(cur)
    return res


import os
SORTCOL = int(os.environ.get("NCU_SORT", "1"))


def main():
    rep, regex, idx, cubin = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    name, hdr, rows = sass_rows(rep, regex, idx)
    func = sys.argv[6] if len(sys.argv) > 6 else re.sub(r"[^A-Za-z0-9_]", "", regex)
    lm = line_map(cubin, func)
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    agg = collections.defaultdict(lambda: [0, 0])
    rows = [r for r in rows if len(r) > ci and (r[ci] or "0").replace(".", "").isdigit()]
    if len(rows) != len(lm):
        print("warning: %d SASS rows in the report vs %d instructions in the cubin" % (len(rows), len(lm)))
    for r, l in zip(rows, lm):
        a = agg[l]
        a[0] += int(r[ci] or 0)
        a[1] += int(r[cs] or 0)
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print("kernel:", name, "| warp instructions:", ti, "| stall samples:", ts)
    src = {}
    for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][SORTCOL] if kv[0] else 0)[:top]:
        if f not in src:
            try:
                import glob
                path = glob.glob("/root/repo/mm_distillnet_b200/csrc/" + f)[0]
                src[f] = open(path).read().splitlines()
            except Exception:
                src[f] = []
        text = src[f][n - 1].strip()[:80] if 0 < n <= len(src[f]) else ""
        print("%-20s %4d inst %9d (%5.1f%%) samples %6d (%5.1f%%) | %s" % (f, n, a[0], 100.0 * a[0] / ti, a[1], 100.0 * a[1] / ts, text))


if __name__ == "__main__":
    main()
