#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1
tail -n 1 gpurun_out/bench_2gpu.log | cut -c1-600
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu.log 2>&1
tail -n 1 gpurun_out/bench_1gpu.log | cut -c1-900
