#!/bin/bash
# full GPU suite + smoke + the default bench line + head / focal benches (evidence at HEAD)
mkdir -p gpurun_out
T=r2m
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
for args in "--batch 16" "--batch 32" "--batch 16 --f32" "--batch 16 --boxes 64"; do
  timeout 300 python tools/focal_bench.py $args 2>&1 | tail -1
done > gpurun_out/r2_focal_bench.jsonl
for args in "--batch 16 --classes 1" "--batch 16 --classes 20" "--batch 32 --classes 20" "--batch 16 --classes 20 --eval" "--batch 16 --classes 20 --f32"; do
  timeout 300 python tools/heads_bench.py $args 2>&1 | tail -2
done > gpurun_out/r2_heads_bench.jsonl
tail -4 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_bench.json | cut -c1-600; cat gpurun_out/r2_focal_bench.jsonl | cut -c1-420
