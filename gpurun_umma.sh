#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -x -q 2>&1 | tail -40 > gpurun_out/pytest_umma.log
tail -25 gpurun_out/pytest_umma.log
