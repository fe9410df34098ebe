#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --model icm --steps 5 --warmup 3 > gpurun_out/bench_icm.log 2>&1; tail -n 1 gpurun_out/bench_icm.log | cut -c1-1500
timeout 600 python bench.py --model disagreement --steps 5 --warmup 3 > gpurun_out/bench_dis.log 2>&1; tail -n 1 gpurun_out/bench_dis.log | cut -c1-1500
