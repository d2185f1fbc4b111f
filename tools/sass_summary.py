#!/usr/bin/env python
"""Opcode histogram per kernel of the shipped library (cuobjdump -sass), the evidence for what the kernels are made of:
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (non-tensor TMA), UTMALDG / UTMASTG = tensor-map TMA,
SYNCS = mbarrier ops, FFMA2 = packed fp32 FMA, MUFU = special function unit, HMMA = legacy mma.sync (none expected).
    python tools/sass_summary.py [path/to/libmmd_b200.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mm_distillnet_b200", "libmmd_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FFMA2", "FFMA", "FMUL2", "FADD2", "MUFU", "F2FP",
        "LDS", "STS", "LDG", "STG", "RED", "ATOMG", "BAR", "DFMA", "DADD", "DMUL"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        funcs[cur]["_total"] += 1
        for k in KEYS:
            if op == k or (k in ("LDG", "STG", "LDS", "STS", "BAR", "MUFU", "RED", "ATOMG") and op.startswith(k)):
                funcs[cur][k] += 1
                break
demangle = subprocess.run(["cu++filt"] + list(funcs), capture_output=True, text=True)
names = demangle.stdout.splitlines() if demangle.returncode == 0 else list(funcs)
print("libmmd_b200.so: SASS opcode counts per kernel (static instruction counts, sm_100a)\n")
print("%-74s %6s " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEYS))
tot = collections.Counter()
for (f, c), n in zip(funcs.items(), names):
    n = n.replace("mmd::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
    n = n[:n.rfind(">(") + 1] if ">(" in n else re.sub(r"\(.*", "", n)
    print("%-74s %6d " % (n[:74], c["_total"]) + " ".join("%7d" % c[k] for k in KEYS))
    tot.update(c)
print("%-74s %6d " % ("TOTAL (%d kernels)" % len(funcs), tot["_total"]) + " ".join("%7d" % tot[k] for k in KEYS))
