#!/bin/bash
mkdir -p gpurun_out
T=r2c
timeout 300 python tools/chain_probe.py 16 > gpurun_out/${T}_probe_default.log 2>&1
MMD_CHAIN_NOCOOP=1 timeout 300 python tools/chain_probe.py 16 > gpurun_out/${T}_probe_nocoop.log 2>&1
MMD_CHAIN_PRE_MAX_HW=144 timeout 300 python tools/chain_probe.py 16 > gpurun_out/${T}_probe_nop5pre.log 2>&1
MMD_CHAIN_NOCOOP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --batch 16 --no-cfg2 > gpurun_out/${T}_bench_nocoop.json 2> gpurun_out/${T}_bench_nocoop.err
MMD_CHAIN_PRE_MAX_HW=144 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --batch 16 --no-cfg2 > gpurun_out/${T}_bench_nop5pre.json 2> gpurun_out/${T}_bench_nop5pre.err
MMD_CHAIN_NOCOOP=1 MMD_CHAIN_PRE_MAX_HW=144 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --batch 16 --no-cfg2 > gpurun_out/${T}_bench_both.json 2> gpurun_out/${T}_bench_both.err
tail -4 gpurun_out/${T}_probe_default.log
