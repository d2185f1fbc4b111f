#!/bin/bash
# quick A/B: bf16 stack parity tests + the bench line (no CPU legs)
mkdir -p gpurun_out
T=${1:-r2q}
timeout 900 python -m pytest tests/test_gpu_bifpn.py -m gpu -q -x -k "bf16 or golden" 2>&1 | tail -4 > gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "b16", d.get("cfg2_b16",{}).get("value"))
for k,v in sorted(d["roofline"]["all_kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:12]:
    print("%-20s %.3f ms  frac %s" % (k, v["ms_per_step"], v["frac"]))
PY
