#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/layer_bench.py --n 16384 --reps 1 --only conv1_wgrad > gpurun_out/plain3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_conv1 -s 1 -c 1 -o gpurun_out/prof_conv1w python tools/layer_bench.py --n 16384 --reps 1 --only conv1_wgrad > gpurun_out/ncu3.log 2>&1
cat gpurun_out/plain3.log | tail -1
for o in "conv1" "conv1_wgrad"; do timeout 100 python tools/layer_bench.py --reps 2 --only $o | tail -2; done
