#!/bin/bash
# round 2, final state: tests, smoke, bench lines (ours + reference arm), widening-row benches, launch list of one step
mkdir -p gpurun_out
T=r2y
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
for args in "--batch 16" "--batch 32" "--batch 16 --f32"; do timeout 300 python tools/focal_bench.py $args 2>&1 | tail -1; done > gpurun_out/${T}_focal_bench.jsonl
for args in "--batch 16 --cpu" "--batch 32" "--batch 16 --f32"; do timeout 300 python tests/pseudo_bench.py $args 2>&1 | tail -1; done > gpurun_out/${T}_pseudo_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${T}_launches_b32.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph --no-cfg2 --min-seconds 0 --batch 32 > gpurun_out/${T}_launches_b32.log 2>&1
tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-300 gpurun_out/${T}_bench.json; cat gpurun_out/${T}_focal_bench.jsonl | cut -c1-300
