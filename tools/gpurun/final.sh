#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; tail -n 1 gpurun_out/bench_full.log | cut -c1-700
timeout 900 python bench.py --model ppo_icm --steps 5 --warmup 3 > gpurun_out/bench_ppo_icm.log 2>&1; tail -n 1 gpurun_out/bench_ppo_icm.log | cut -c1-260
timeout 600 python bench.py --model icm --steps 5 --warmup 3 > gpurun_out/bench_icm.log 2>&1; tail -n 1 gpurun_out/bench_icm.log | cut -c1-260
timeout 600 python bench.py --model disagreement --steps 5 --warmup 3 > gpurun_out/bench_dis.log 2>&1; tail -n 1 gpurun_out/bench_dis.log | cut -c1-260
timeout 600 python bench.py --model ndigo --B 4096 --steps 5 --warmup 3 > gpurun_out/bench_ndigo.log 2>&1; tail -n 1 gpurun_out/bench_ndigo.log | cut -c1-260
