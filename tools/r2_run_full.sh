#!/bin/bash
mkdir -p gpurun_out
for B in 16 32; do timeout 600 python tools/full_step_bench.py --batch $B 2>&1 | tail -3; done > gpurun_out/r2_full_step_bench.jsonl
cat gpurun_out/r2_full_step_bench.jsonl | cut -c1-900
