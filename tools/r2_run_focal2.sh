#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_focal.py tests/test_gpu_full_step.py tests/test_gpu_pseudo.py -q -m gpu --no-header -x 2>&1 | tail -3 > gpurun_out/r2f2_tests.log
for args in "--batch 16" "--batch 32" "--batch 16 --boxes 64" "--batch 16 --boxes 256"; do timeout 300 python tools/focal_bench.py $args 2>&1 | tail -1; done > gpurun_out/r2f2_focal_bench.jsonl
cat gpurun_out/r2f2_tests.log; cut -c1-260 gpurun_out/r2f2_focal_bench.jsonl
