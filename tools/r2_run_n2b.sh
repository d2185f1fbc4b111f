#!/bin/bash
mkdir -p gpurun_out
T=r2n2
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_bench.err; cut -c1-200 gpurun_out/${T}_bench.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("grad_checksum"))
PY
