#!/bin/bash
# round 2, GPU run G (2 GPUs): N = 2 NCCL gradient-parity test, bench at N = 2
mkdir -p gpurun_out
T=r2g
nvidia-smi -L > gpurun_out/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_bifpn.py -m gpu -q -k "nccl or product or eval_mode" > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --min-seconds 1 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "rc=$?" >> gpurun_out/${T}_bench_n2.err
tail -5 gpurun_out/${T}_tests.log; tail -c 400 gpurun_out/${T}_bench_n2.json; tail -3 gpurun_out/${T}_bench_n2.err
