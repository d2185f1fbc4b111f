#!/bin/bash
mkdir -p gpurun_out
T=r2o
timeout 900 python -m pytest tests/test_gpu_optim.py tests/test_gpu_bifpn.py -m gpu -q -x -k "optim or adam or replay or distill or graph" 2>&1 | tail -6 > gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "b16", d.get("cfg2_b16",{}).get("value"), "launches/step", d["gpu_launches_per_step"], d["grad_checksum"])
print({k:v for k,v in d["roofline"]["all_kernels"].items() if k in ("adam",)})
PY
