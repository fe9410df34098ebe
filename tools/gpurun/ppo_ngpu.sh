#!/bin/bash
# usage: ppo_ngpu.sh N [extra bench flags]
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --model ppo_icm --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_ppo_icm_${N}gpu.log 2>&1
tail -n 1 gpurun_out/bench_ppo_icm_${N}gpu.log | cut -c1-1800
