#!/bin/bash
mkdir -p gpurun_out
T=r2l
for S in 1.5 3.0; do
MMD_POOL_TILED_TRAIN_SHARE=$S timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 --no-cfg2 > gpurun_out/${T}_share$S.json 2> gpurun_out/${T}_share$S.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
