#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/policy_breakdown.py > gpurun_out/policy_breakdown.txt 2>&1; cat gpurun_out/policy_breakdown.txt | tail -20
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
