#!/bin/bash
# round 2, evidence pack at HEAD: tests, diagnostics, bench lines, launch lists with DRAM bytes, ncu --set full captures,
# compute-sanitizer.  Everything lands in gpurun_out/r2f_* (copied / summarised into profiles/ afterwards).
mkdir -p gpurun_out
T=r2z
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 900 python tests/run_gpu_diag.py > gpurun_out/${T}_diag.log 2>&1; cp gpurun_out/diag.json gpurun_out/${T}_diag.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
MMD_SPLIT_STUDENT=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench_split.json 2> gpurun_out/${T}_bench_split.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 300 python bench.py --dtype f32 --steps 10 --warmup 3 --no-cpu --min-seconds 1 --no-cfg2 --batch 16 > gpurun_out/${T}_bench_f32_b16.json 2> gpurun_out/${T}_bench_f32.err
timeout 300 python tools/eval_sweep.py > gpurun_out/${T}_eval_sweep.txt 2>&1
timeout 300 python tests/torch_gpu_comparator.py --batch 32 --steps 5 > gpurun_out/${T}_torch_gpu_b32.txt 2>&1
for B in 16 32; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${T}_launches_b${B}.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph --no-cfg2 --min-seconds 0 --batch $B > gpurun_out/${T}_launches_b${B}.log 2>&1
done
# full counters + source: one cell of the forward (8 node launches + 4 poolfuse), projections, MTA; one cell of the backward
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:'node_fwd_v4_kernel|poolfuse' -s 152 -c 12 -o /tmp/${T}_full_fwd python tools/step_probe.py 16 > gpurun_out/${T}_ncu_fwd.log 2>&1
timeout 900 $NCU -k regex:'proj_fwd_tma|proj_bwd4|mta_pool_c112|mta_bwd_c112|mta_level|bnapply_same|slot_same' -s 17 -c 17 -o /tmp/${T}_full_misc python tools/step_probe.py 16 > gpurun_out/${T}_ncu_misc.log 2>&1
timeout 900 $NCU -k regex:'node_bwd_a4|node_bwd_b4' -s 96 -c 16 -o /tmp/${T}_full_bwd python tools/step_probe.py 16 > gpurun_out/${T}_ncu_bwd.log 2>&1
# (the .ncu-rep files stay on the box: gpurun_out/ is limited to 64 MiB; raw counters and the source pages of the main kernels come back as CSV)
for P in fwd misc bwd; do ncu -i /tmp/${T}_full_${P}.ncu-rep --page raw --csv > gpurun_out/${T}_full_${P}_raw.csv 2>/dev/null; done
ncu -i /tmp/${T}_full_fwd.ncu-rep --page source --csv --kernel-id '::regex:node_fwd_v4_kernel<\(int\)16:2' > gpurun_out/${T}_src_node_fwd_p3.csv 2>/dev/null
ncu -i /tmp/${T}_full_bwd.ncu-rep --page source --csv --kernel-id '::regex:node_bwd_a4_kernel<\(int\)16:1' > gpurun_out/${T}_src_node_bwd_a.csv 2>/dev/null
ncu -i /tmp/${T}_full_bwd.ncu-rep --page source --csv --kernel-id '::regex:node_bwd_b4_kernel<\(int\)16:1' > gpurun_out/${T}_src_node_bwd_b.csv 2>/dev/null
ncu -i /tmp/${T}_full_misc.ncu-rep --page source --csv --kernel-id '::regex:proj_fwd_tma:1' > gpurun_out/${T}_src_proj_tma.csv 2>/dev/null
ls -la /tmp/${T}_full_*.ncu-rep > gpurun_out/${T}_ncu_rep_sizes.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/${T}_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/${T}_racecheck.log 2>&1
ls -la gpurun_out | grep ${T}; tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_memcheck.log; tail -2 gpurun_out/${T}_racecheck.log
