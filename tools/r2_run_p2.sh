#!/bin/bash
# pseudo-label evidence: graph-replay test, sanitizer passes over all kernel families incl. pseudo, ncu --set full of the two passes, smoke
mkdir -p gpurun_out
T=r2p
timeout 600 python -m pytest tests/test_gpu_pseudo.py -q -m gpu --no-header -rf -x 2>&1 | tail -15 > gpurun_out/${T}_tests.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/${T}_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/${T}_racecheck.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:'pl_' -s 15 -c 5 -o /tmp/${T}_pseudo python tests/pseudo_bench.py --batch 16 --steps 1 --warmup 2 > gpurun_out/${T}_ncu.log 2>&1
ncu -i /tmp/${T}_pseudo.ncu-rep --page raw --csv > gpurun_out/${T}_pseudo_raw.csv 2>/dev/null
cat gpurun_out/${T}_tests.log; tail -4 gpurun_out/${T}_memcheck.log; tail -4 gpurun_out/${T}_racecheck.log; tail -3 gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_ncu.log; ls -la gpurun_out | grep ${T}
