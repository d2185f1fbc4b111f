#!/bin/bash
# A/B of an environment switch: bf16 parity tests with the switch on, then the bench line off / on
mkdir -p gpurun_out
SW="$1"; out=gpurun_out/r2_ab_${SW%%=*}.txt; : > $out
env $SW timeout 900 python -m pytest tests/test_gpu_bifpn.py -m gpu -q -x -k "bf16 or golden" 2>&1 | tail -2 >> $out
run() { echo -n "$* : " >> $out; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-cfg2 --min-seconds 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['all_kernels']
print(round(d['value'],1), round(d['ms_per_step'],4), {n: round(k[n]['ms_per_step'],3) for n in ('node_bwd_a','node_bwd_b','node_fwd') if n in k})" >> $out; }
run MMD_NOP=1
run $SW
cat $out
