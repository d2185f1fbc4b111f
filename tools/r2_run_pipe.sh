#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pseudo.py -q -m gpu --no-header -x 2>&1 | tail -3 > gpurun_out/r2pp_tests.log
for m in lockstep pipe_bwd pipe_all; do timeout 300 python tools/pipe_probe.py --mode $m 2>&1 | tail -2; done > gpurun_out/r2_pipe_probe.jsonl
cat gpurun_out/r2pp_tests.log gpurun_out/r2_pipe_probe.jsonl
