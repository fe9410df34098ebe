#!/bin/bash
mkdir -p gpurun_out
M=${1:-icm}
timeout 300 python bench.py --model $M --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$M.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$M.csv python bench.py --model $M --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$M.log 2>&1
tail -n 2 gpurun_out/ncu_$M.log | cut -c1-300
