#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1
cat gpurun_out/layer_bench.log
timeout 200 python tools/layer_bench.py --n 32768 --reps 1 --only conv2_fwd > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 2 -c 1 -o gpurun_out/prof_conv2 python tools/layer_bench.py --n 32768 --reps 1 --only conv2_fwd > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
