#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pseudo.py tests/test_gpu_focal.py -q -m gpu --no-header -rf -x 2>&1 | tail -60 > gpurun_out/r2_pseudo_tests.log
for args in "--batch 16 --cpu" "--batch 32" "--batch 16 --f32"; do
  timeout 300 python tests/pseudo_bench.py $args 2>&1 | tail -3
done > gpurun_out/r2_pseudo_bench.jsonl
cat gpurun_out/r2_pseudo_tests.log gpurun_out/r2_pseudo_bench.jsonl
