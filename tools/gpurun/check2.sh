#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/policy_breakdown.py > gpurun_out/policy_breakdown.txt 2>&1; tail -n 13 gpurun_out/policy_breakdown.txt
timeout 900 python bench.py --model ppo_icm --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ppo_icm_short.log 2>&1; tail -n 1 gpurun_out/bench_ppo_icm_short.log | cut -c1-260
timeout 600 python bench.py --model icm --steps 5 --warmup 3 > gpurun_out/bench_icm.log 2>&1; tail -n 1 gpurun_out/bench_icm.log | cut -c1-260
timeout 600 python bench.py --model disagreement --steps 5 --warmup 3 > gpurun_out/bench_dis.log 2>&1; tail -n 1 gpurun_out/bench_dis.log | cut -c1-260
