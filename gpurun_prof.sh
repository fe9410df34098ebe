#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/layer_bench.py --reps 1 --only conv2_fwd > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_window -s 2 -c 1 -o gpurun_out/prof_conv2_full python tools/layer_bench.py --reps 1 --only conv2_fwd > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
