#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_full.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --B 128 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --B 128 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_full.log; tail -2 gpurun_out/ncu.log
