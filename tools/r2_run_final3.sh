#!/bin/bash
# round 2, final state: tests, smoke, sanitizer, bench lines (ours + reference arm), launch list of one step
mkdir -p gpurun_out
T=r2x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/${T}_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/${T}_racecheck.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${T}_launches_b32.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph --no-cfg2 --min-seconds 0 --batch 32 > gpurun_out/${T}_launches_b32.log 2>&1
tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_memcheck.log; tail -2 gpurun_out/${T}_racecheck.log; cut -c1-250 gpurun_out/${T}_bench.json; cut -c1-250 gpurun_out/${T}_bench_reference.json
