#!/bin/bash
mkdir -p gpurun_out
T=r2db
timeout 900 python -m pytest tests/test_gpu_bifpn.py -m gpu -q -x -k "replay or distill or graph" 2>&1 | tail -6 > gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "b16", d.get("cfg2_b16",{}).get("value"), d.get("cfg2_b16",{}).get("e2e"))
PY
