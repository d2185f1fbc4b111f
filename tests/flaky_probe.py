#!/usr/bin/env python
"""Run the fp32 stack-vs-oracle cases several times and print the run-to-run spread of every gradient metric
(tolerance calibration for tests/test_gpu_bifpn.py::test_stack_vs_oracle_fp32)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_cases as G  # noqa: E402

CASES = [((3, True, 2, 48), {"fw_mode": "mixed", "channels_last": True}, 6), ((5, True, 2, 96), {}, 3)]
for args, kw, reps in CASES:
    runs = [G.random_stack_case(*args, **kw) for _ in range(reps)]
    print("case", args, kw)
    for k in sorted(runs[0]):
        if k.startswith("worstname_"):
            print("   %-22s %s" % (k, " | ".join(str(r.get(k)) for r in runs[:3])))
        elif k.startswith(("pgrad_", "grad_in", "zero_grad", "ref32_pgrad", "ref32_grad", "train_", "eval_")):
            print("   %-22s %s" % (k, " ".join("%.2e" % r.get(k, float("nan")) for r in runs)))
