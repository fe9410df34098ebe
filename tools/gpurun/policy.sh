#!/bin/bash
# policy / PPO parity tests, stage breakdown and the configs[3] bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py -q --tb=short > gpurun_out/policy_tests.log 2>&1; tail -n 40 gpurun_out/policy_tests.log
timeout 600 python tools/policy_breakdown.py > gpurun_out/policy_breakdown.txt 2>&1; tail -n 20 gpurun_out/policy_breakdown.txt
timeout 600 python bench.py --model ppo_icm --steps 5 --warmup 3 > gpurun_out/bench_ppo_icm.log 2>&1; tail -n 2 gpurun_out/bench_ppo_icm.log | cut -c1-1500
