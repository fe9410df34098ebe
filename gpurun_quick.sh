#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/layer_bench.py 2>&1 | tee gpurun_out/layer_bench.log
