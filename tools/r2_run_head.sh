#!/bin/bash
# HEAD check: full GPU test suite, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
T=r2h
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -4 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-600 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench_reference.json
