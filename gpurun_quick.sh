#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py tests/test_gpu_rnd.py tests/test_gpu_layers.py -x -q 2>&1 | tail -4
timeout 300 python tools/layer_bench.py --only wgrad 2>&1 | tee gpurun_out/layer_bench.log
