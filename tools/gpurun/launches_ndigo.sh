#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --model ndigo --B 4096 --steps 1 --warmup 1 > gpurun_out/plain_ndigo.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_ndigo.csv python bench.py --model ndigo --B 4096 --steps 1 --warmup 1 > gpurun_out/ncu_ndigo.log 2>&1
tail -n 1 gpurun_out/ncu_ndigo.log | cut -c1-200
