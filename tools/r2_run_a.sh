#!/bin/bash
# round 2, GPU run A: parity tests, full diagnostics (all metrics, nothing asserted), the new bench line
mkdir -p gpurun_out
T=r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 900 python tests/run_gpu_diag.py stack > gpurun_out/${T}_diag.log 2>&1; cp gpurun_out/diag.json gpurun_out/${T}_diag.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
tail -3 gpurun_out/${T}_tests.log; tail -c 600 gpurun_out/${T}_bench.json
