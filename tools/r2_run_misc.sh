#!/bin/bash
mkdir -p gpurun_out
T=r2w
timeout 300 python bench.py --dtype f32 --steps 10 --warmup 3 --no-cpu --min-seconds 1 --no-cfg2 --batch 16 > gpurun_out/${T}_bench_f32_b16.json 2> gpurun_out/${T}_bench_f32.err; echo "rc=$?" >> gpurun_out/${T}_bench_f32.err
timeout 300 python tools/eval_sweep.py > gpurun_out/${T}_eval_sweep.txt 2>&1
tail -2 gpurun_out/${T}_bench_f32.err; cut -c1-200 gpurun_out/${T}_bench_f32_b16.json; tail -12 gpurun_out/${T}_eval_sweep.txt
