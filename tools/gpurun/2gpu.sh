#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1
tail -n 1 gpurun_out/bench_2gpu.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --model disagreement > gpurun_out/bench_dis_2gpu.log 2>&1
tail -n 1 gpurun_out/bench_dis_2gpu.log | cut -c1-400
