#!/bin/bash
# round 2, GPU run B: forward chain kernel — parity tests, diagnostics, bench with and without the chain
mkdir -p gpurun_out
T=r2b
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 900 python tests/run_gpu_diag.py stack > gpurun_out/${T}_diag.log 2>&1; cp gpurun_out/diag.json gpurun_out/${T}_diag.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_bench.err
MMD_NO_CHAIN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --min-seconds 1 > gpurun_out/${T}_bench_nochain.json 2> gpurun_out/${T}_bench_nochain.err
tail -5 gpurun_out/${T}_tests.log; tail -c 300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
