#!/bin/bash
mkdir -p gpurun_out
T=r2s
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k node_fwd_v4_kernel -s 42 -c 2 -o gpurun_out/${T}_fwd python tools/step_probe.py 16 > gpurun_out/${T}_ncu_fwd.log 2>&1
timeout 600 $NCU -k node_bwd_a4_kernel -s 43 -c 2 -o gpurun_out/${T}_bwd_a python tools/step_probe.py 16 > gpurun_out/${T}_ncu_bwd_a.log 2>&1
timeout 600 $NCU -k node_bwd_b4_kernel -s 44 -c 1 -o gpurun_out/${T}_bwd_b python tools/step_probe.py 16 > gpurun_out/${T}_ncu_bwd_b.log 2>&1
ls -la gpurun_out/
