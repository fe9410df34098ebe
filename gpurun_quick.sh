#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -k conv1 -x -q 2>&1 | tail -12
timeout 300 python tools/layer_bench.py --only conv1 2>&1 | tee gpurun_out/layer_bench.log
