#!/bin/bash
mkdir -p gpurun_out
T=r2k
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_tests.log
