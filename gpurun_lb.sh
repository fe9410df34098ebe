#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_layers.py tests/test_gpu_rnd.py -x -q 2>&1 | tail -3
timeout 300 python tools/layer_bench.py --only "${1:-}" 2>&1 | tee gpurun_out/layer_bench.log
