#!/bin/bash
# environment-knob sweep at B = 32 (CTA split between the training network and the frozen ones in lockstep launches)
mkdir -p gpurun_out
out=gpurun_out/r2_knob_sweep.txt; : > $out
run() { echo -n "$* : " >> $out; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-cfg2 --min-seconds 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4))" >> $out; }
run MMD_NOP=1
run MMD_FWD_TRAIN_SHARE=0.8
run MMD_FWD_TRAIN_SHARE=1.25
run MMD_FWD_TRAIN_SHARE=1.5
run MMD_POOL_TILED_TRAIN_SHARE=1.5
run MMD_POOL_TILED_TRAIN_SHARE=3.0
run MMD_POOL_TRAIN_SHARE=2.0
run MMD_POOL_TRAIN_SHARE=4.0
run MMD_NOP=2
cat $out
