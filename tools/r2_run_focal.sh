#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_focal.py -q -m gpu --no-header -rf 2>&1 | tail -40 > gpurun_out/r2_focal_tests.log
for args in "--batch 16" "--batch 32" "--batch 16 --f32" "--batch 16 --boxes 64"; do
  timeout 300 python tools/focal_bench.py $args 2>&1 | tail -2
done > gpurun_out/r2_focal_bench.jsonl
cat gpurun_out/r2_focal_tests.log gpurun_out/r2_focal_bench.jsonl
