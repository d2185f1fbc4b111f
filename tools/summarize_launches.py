#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list per
(kernel, grid) into a markdown table; with the DRAM metrics present it also prints DRAM bytes per launch and can write
the per-kernel-kind traffic file bench.py reads (profiles/traffic.json).

    python tools/summarize_launches.py gpurun_out/launches.csv [--last N] [--traffic profiles/traffic.json bf16]

The per-launch times of such a pass are cold-cache and serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import json
import os
import re
import sys

KINDS = [("node_fwd<16,8>", r"node_fwd_v4_kernel<\(?(int\))?16, \(?(int\))?8[,>]"), ("node_bwd_a<16,8>", r"node_bwd_a4_kernel<\(?(int\))?16, \(?(int\))?8[,>]"),
         ("node_bwd_b<16,8>", r"node_bwd_b4_kernel<\(?(int\))?16, \(?(int\))?8"), ("node_fwd", r"node_fwd"), ("poolfuse", r"poolfuse"), ("proj_fwd", r"proj_fwd"), ("bnapply", r"bnapply"),
         ("node_bwd_a", r"node_bwd_a"), ("node_bwd_b", r"node_bwd_b"), ("proj_bwd", r"proj_bwd"), ("pull", r"pull_kernel"),
         ("slot", r"slot_(same|group_|kernel)"), ("mta_pool", r"mta_pool"), ("mta_level", r"mta_level"), ("mta_bwd", r"mta_bwd"),
         ("mta_finish", r"mta_finish"), ("prep", r"prep_kernel"), ("bn_finalize", r"bn_finalize"), ("fwgrad", r"fwgrad")]


def to_us(v, unit):
    unit = (unit or "ns").lower()
    if unit.startswith("ns"):
        return v / 1e3
    if unit.startswith("us"):
        return v
    if unit.startswith("ms"):
        return v * 1e3
    if unit.startswith("s"):
        return v * 1e6
    return v / 1e3


def to_bytes(v, unit):
    unit = (unit or "byte").lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    return v * mult.get(unit, 1)


def main(argv):
    path = argv[1]
    last = int(argv[argv.index("--last") + 1]) if "--last" in argv else None
    rows = list(csv.reader(ln for ln in open(path) if not ln.startswith("==")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ii, ki, vi, gi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    mi, ui = hdr.index("Metric Name"), hdr.index("Metric Unit")
    launches = collections.OrderedDict()   # id -> dict
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
        L = launches.setdefault(r[ii], {"name": name[:70], "grid": r[gi], "us": 0.0, "rd": None, "wr": None})
        m = r[mi]
        if m == "gpu__time_duration.sum":
            L["us"] = to_us(v, r[ui])
        elif m == "dram__bytes_read.sum":
            L["rd"] = to_bytes(v, r[ui])
        elif m == "dram__bytes_write.sum":
            L["wr"] = to_bytes(v, r[ui])
    seq = list(launches.values())
    if last:
        seq = seq[-last:]
    has_dram = any(L["rd"] is not None for L in seq)
    agg = collections.defaultdict(list)
    for L in seq:
        agg[(L["name"], L["grid"])].append(L)
    tot = sum(L["us"] for L in seq) or 1.0
    if has_dram:
        print("| kernel | grid | launches | total us | avg us | share | DRAM MB / launch (rd+wr) |")
        print("|---|---|---:|---:|---:|---:|---:|")
    else:
        print("| kernel | grid | launches | total us | avg us | share |")
        print("|---|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(L["us"] for L in kv[1])):
        t = sum(L["us"] for L in v)
        line = "| `%s` | %s | %d | %.1f | %.1f | %.3f |" % (k[0], k[1], len(v), t, t / len(v), t / tot)
        if has_dram:
            line += " %.2f |" % (sum((L["rd"] or 0) + (L["wr"] or 0) for L in v) / len(v) / 1e6)
        print(line)
    print()
    print("total device time in the list: %.1f us over %d launches" % (tot, len(seq)))
    if "--traffic" in argv:
        out, dtype = argv[argv.index("--traffic") + 1], argv[argv.index("--traffic") + 2]
        kinds = {}
        def kind_of(name):   # first matching kind wins (the <16,8> instantiations are kinds of their own)
            for kind, pat in KINDS:
                if re.search(pat, name):
                    return kind
            return None
        for kind, _ in KINDS:
            sel = [L for L in seq if kind_of(L["name"]) == kind and L["rd"] is not None]
            if sel:
                kinds[kind] = sum(L["rd"] + (L["wr"] or 0) for L in sel) / len(sel)
        cur = {}
        if os.path.exists(out):
            cur = json.load(open(out))
        cur[dtype] = kinds
        cur.setdefault("_note", "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, averaged per kernel "
                                "kind over the launches of the profiled window; from tools/summarize_launches.py")
        json.dump(cur, open(out, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv)
