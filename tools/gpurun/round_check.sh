#!/bin/bash
# full GPU suite + the bench lines of every mode (short runs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --model icm --steps 5 --warmup 3 > gpurun_out/bench_icm.log 2>&1; tail -n 1 gpurun_out/bench_icm.log | cut -c1-420
timeout 600 python bench.py --model disagreement --steps 5 --warmup 3 > gpurun_out/bench_dis.log 2>&1; tail -n 1 gpurun_out/bench_dis.log | cut -c1-420
timeout 600 python bench.py --model ppo_icm --steps 5 --warmup 3 > gpurun_out/bench_ppo_icm.log 2>&1; tail -n 1 gpurun_out/bench_ppo_icm.log | cut -c1-420
