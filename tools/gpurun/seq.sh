#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ndigo.py tests/test_gpu_policy.py -q --tb=short > gpurun_out/seq_tests.log 2>&1; tail -n 30 gpurun_out/seq_tests.log
timeout 600 python bench.py --model ndigo --B 4096 --steps 5 --warmup 3 > gpurun_out/bench_ndigo.log 2>&1; tail -n 1 gpurun_out/bench_ndigo.log | cut -c1-500
timeout 600 python tools/policy_breakdown.py > gpurun_out/policy_breakdown.txt 2>&1; tail -n 14 gpurun_out/policy_breakdown.txt
timeout 900 python bench.py --model ppo_icm --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ppo_icm_short.log 2>&1; tail -n 1 gpurun_out/bench_ppo_icm_short.log | cut -c1-300
