#!/bin/bash
# configs[3] (ICM + LSTM PPO): launch list of one iteration under ncu, after the plain run exited 0
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py -q --tb=short 2>&1 | tail -3
timeout 600 python bench.py --model ppo_icm --steps 1 --warmup 1 > gpurun_out/plain_ppo_icm.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_ppo_icm.csv python bench.py --model ppo_icm --steps 1 --warmup 1 > gpurun_out/ncu_ppo_icm.log 2>&1
tail -n 1 gpurun_out/plain_ppo_icm.log | cut -c1-300
wc -l gpurun_out/launches_ppo_icm.csv
