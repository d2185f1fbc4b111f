#!/bin/bash
mkdir -p gpurun_out
T=r2j
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
MMD_NO_POOL_TILED=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --no-cfg2 > gpurun_out/${T}_bench_notiled.json 2> gpurun_out/${T}_bench_notiled.err
MMD_POOL_TILED_TRAIN_SHARE=1.0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --no-cfg2 > gpurun_out/${T}_bench_share1.json 2> gpurun_out/${T}_bench_share1.err
MMD_POOL_TILED_TRAIN_SHARE=2.5 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --no-cfg2 > gpurun_out/${T}_bench_share25.json 2> gpurun_out/${T}_bench_share25.err
tail -5 gpurun_out/${T}_tests.log
