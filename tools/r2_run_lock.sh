#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_heads.py -q -m gpu --no-header -x -k "lockstep" 2>&1 | grep -E "^E  |passed|failed|Error" | head -8 > gpurun_out/r2l_tests.log
for args in "--batch 32" "--batch 32 --lockstep" "--batch 16 --lockstep"; do timeout 600 python tools/full_step_bench.py $args 2>&1 | tail -1; done > gpurun_out/r2_full_step_lockstep.jsonl
cat gpurun_out/r2l_tests.log; cut -c200-760 gpurun_out/r2_full_step_lockstep.jsonl
