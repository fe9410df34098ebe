#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ndigo.py tests/test_gpu_icm.py -x -q 2>&1 | grep -E "^E|passed|failed|Error" | head -12
