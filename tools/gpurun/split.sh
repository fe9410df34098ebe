#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_ndigo.py -q --tb=short 2>&1 | grep -v Warn | tail -12
timeout 300 python tools/policy_breakdown.py 2>&1 | grep -E "time loop|forward \(all\)|backward \(all\)"
timeout 300 python bench.py --model ndigo --B 4096 --steps 5 --warmup 3 2>&1 | tail -n 1 | cut -c1-230
