#!/bin/bash
# heads: parity tests, full GPU suite, device-timed head steps, launch list of one head step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_heads.py -q -m gpu --no-header -rf 2>&1 | tail -40 > gpurun_out/r2_heads_tests.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_heads.py 2>&1 | tail -8 > gpurun_out/r2_heads_suite.log
for args in "--batch 16 --classes 1" "--batch 16 --classes 20" "--batch 32 --classes 1" "--batch 16 --classes 1 --eval" "--batch 16 --classes 1 --f32"; do
  timeout 300 python tools/heads_bench.py $args 2>&1 | tail -2
done > gpurun_out/r2_heads_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_heads_launches.csv \
  python tools/heads_bench.py --batch 16 --classes 1 --steps 1 --warmup 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_heads_launches.csv > gpurun_out/r2_heads_launches.md 2>&1
tail -30 gpurun_out/r2_heads_tests.log; cat gpurun_out/r2_heads_suite.log gpurun_out/r2_heads_bench.jsonl; head -40 gpurun_out/r2_heads_launches.md
