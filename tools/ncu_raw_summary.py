#!/usr/bin/env python
"""Markdown table of selected counters from `ncu -i X.ncu-rep --page raw --csv` exports (one row per profiled launch).
    python tools/ncu_raw_summary.py gpurun_out/r2d_fwd_raw.csv [more.csv ...] > profiles/r2_ncu_summary.md"""
import csv
import re
import sys

COLS = [("us", "gpu__time_duration.sum", 1e-3), ("DRAM rd MB", "dram__bytes_read.sum", None), ("DRAM wr MB", "dram__bytes_write.sum", None),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("issue %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1), ("warp inst", "smsp__inst_executed.sum", 1),
        ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1), ("smem conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
        ("tensor %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", 1), ("regs", "launch__registers_per_thread", 1),
        ("st long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1),
        ("st barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1),
        ("st short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1),
        ("st wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1),
        ("st mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", 1),
        ("st math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1),
        ("st sleep", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", 1)]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def unit_scale(unit):
    u = (unit or "").lower()
    return {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, None)


def main(paths):
    print("| kernel | grid | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for path in paths:
        rows = list(csv.reader(ln for ln in open(path) if not ln.startswith("==")))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in data:
            if len(r) < len(hdr):
                continue
            name = r[ix["Kernel Name"]]
            name = re.sub(r"\(.*", "", name.replace("mmd::", "").replace("(int)", "").replace("(bool)", ""))
            cells = []
            for label, key, scale in COLS:
                if key not in ix:
                    hits = [h for h in hdr if h.startswith(key.split(".")[0])]
                    cells.append("n/a")
                    continue
                v = num(r[ix[key]])
                if v is None:
                    cells.append(r[ix[key]])
                    continue
                if scale is None:
                    sc = unit_scale(units[ix[key]])
                    v = v * (sc if sc is not None else 1e-6)
                elif label == "us":
                    u = (units[ix[key]] or "").lower()
                    v = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
                cells.append("%.4g" % v)
            print("| `%s` | %s | " % (name[:60], r[ix["Grid Size"]]) + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1:])
