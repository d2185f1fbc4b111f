#!/bin/bash
# evidence for the widening rows: sanitizer passes incl. heads + focal kernels, ncu --set full of the head glue and focal kernels
mkdir -p gpurun_out
T=r2e
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/${T}_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/${T}_racecheck.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:'focal_kernel' -s 2 -c 2 -o /tmp/${T}_focal python tools/focal_bench.py --batch 16 --steps 1 --warmup 2 > gpurun_out/${T}_ncu_focal.log 2>&1
timeout 600 $NCU -k regex:'head_gather|head_scatter|act_fwd|act_bwd' -s 12 -c 12 -o /tmp/${T}_heads python tools/heads_bench.py --batch 16 --classes 20 --steps 1 --warmup 1 > gpurun_out/${T}_ncu_heads.log 2>&1
for P in focal heads; do ncu -i /tmp/${T}_${P}.ncu-rep --page raw --csv > gpurun_out/${T}_${P}_raw.csv 2>/dev/null; done
tail -3 gpurun_out/${T}_memcheck.log; tail -3 gpurun_out/${T}_racecheck.log; ls -la gpurun_out | grep ${T}
