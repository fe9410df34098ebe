#!/bin/bash
# policy / PPO parity tests + the configs[3] bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py -q --tb=short > gpurun_out/policy_tests.log 2>&1; tail -n 60 gpurun_out/policy_tests.log
timeout 600 python bench.py --model ppo_icm --steps 3 --warmup 3 > gpurun_out/bench_ppo_icm.log 2>&1; tail -n 5 gpurun_out/bench_ppo_icm.log | cut -c1-1500
