#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
T=r2n${N}
nvidia-smi -L > gpurun_out/${T}_smi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --min-seconds 1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
tail -c 300 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
