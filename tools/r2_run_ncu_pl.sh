#!/bin/bash
mkdir -p gpurun_out
T=r2v
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:'pl_|adam_kernel' -s 15 -c 5 -o /tmp/${T}_pseudo python tests/pseudo_bench.py --batch 32 --steps 1 --warmup 2 > gpurun_out/${T}_ncu.log 2>&1
ncu -i /tmp/${T}_pseudo.ncu-rep --page raw --csv > gpurun_out/${T}_pseudo_raw.csv 2>/dev/null
tail -2 gpurun_out/${T}_ncu.log | cut -c1-200; ls -la gpurun_out | grep ${T}
