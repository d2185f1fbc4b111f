#!/bin/bash
mkdir -p gpurun_out
T=r2i
timeout 600 python -m pytest tests/test_gpu_bifpn.py -m gpu -q -x -k "proj_tma" > gpurun_out/${T}_tests_tma.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests_tma.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
MMD_NO_PROJ_TMA=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --no-cfg2 > gpurun_out/${T}_bench_notma.json 2> gpurun_out/${T}_bench_notma.err
tail -15 gpurun_out/${T}_tests_tma.log; tail -3 gpurun_out/${T}_tests.log
