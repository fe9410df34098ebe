#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_icm.py -x -q -k maze 2>&1 | grep -E "^E|assert" | head -12
