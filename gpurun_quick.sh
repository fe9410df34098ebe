#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_returns_hooks.py -x -q 2>&1 | grep -E "^E|passed|failed|Error" | head -12
