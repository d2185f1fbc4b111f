#!/bin/bash
mkdir -p gpurun_out
T=r2e
timeout 600 python -m pytest tests/test_gpu_mta.py -m gpu -q > gpurun_out/${T}_tests_mta.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests_mta.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
MMD_NO_MTA_FAST=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --no-cfg2 > gpurun_out/${T}_bench_nomtafast.json 2> gpurun_out/${T}_bench_nomtafast.err
MMD_DEBUG_SKIP_DW_FLUSH=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench_skipflush.json 2> gpurun_out/${T}_bench_skipflush.err
tail -3 gpurun_out/${T}_tests_mta.log
