#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/layer_bench.py --n 32768 --reps 1 --only conv1_wgrad > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1s_wgrad -s 2 -c 1 -o gpurun_out/prof_conv1w python tools/layer_bench.py --n 32768 --reps 1 --only conv1_wgrad > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
