#!/bin/bash
# round 2, GPU run D: parity tests at HEAD, bf16 diagnostics, ncu evidence pack (full counters per kernel, launch lists
# with DRAM bytes at B=16 and B=32), compute-sanitizer
mkdir -p gpurun_out
T=r2d
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python tests/run_gpu_diag.py bf16 > gpurun_out/${T}_diag.log 2>&1; cp gpurun_out/diag.json gpurun_out/${T}_diag_bf16.json
K='node_fwd_v4_kernel|node_bwd_a4|node_bwd_b4|poolfuse|proj_fwd_tc|proj_bwd4|mta_pool|mta_bwd|mta_level'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 156 -c 68 -f -o /tmp/${T}_fwd python tools/step_probe.py 16 > gpurun_out/${T}_ncu_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 224 -c 90 -f -o /tmp/${T}_bwd python tools/step_probe.py 16 > gpurun_out/${T}_ncu_bwd.log 2>&1
for P in fwd bwd; do
  ncu -i /tmp/${T}_${P}.ncu-rep --page raw --csv > gpurun_out/${T}_${P}_raw.csv 2>/dev/null
done
ncu -i /tmp/${T}_fwd.ncu-rep --page source --csv -k regex:'node_fwd_v4_kernel<16, 8' -c 4 > gpurun_out/${T}_src_node_fwd.csv 2>/dev/null
ncu -i /tmp/${T}_bwd.ncu-rep --page source --csv -k regex:'node_bwd_a4_kernel<16, 8' -c 3 > gpurun_out/${T}_src_node_bwd_a.csv 2>/dev/null
ncu -i /tmp/${T}_bwd.ncu-rep --page source --csv -k regex:'node_bwd_b4_kernel<16, 8' -c 3 > gpurun_out/${T}_src_node_bwd_b.csv 2>/dev/null
for B in 16 32; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${T}_launches_b${B}.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph --no-cfg2 --min-seconds 0 --batch $B > gpurun_out/${T}_launches_b${B}.log 2>&1
done
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/${T}_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/${T}_racecheck.log 2>&1
tail -3 gpurun_out/${T}_tests.log; ls -la gpurun_out | grep ${T}; tail -3 gpurun_out/${T}_memcheck.log; tail -3 gpurun_out/${T}_racecheck.log
