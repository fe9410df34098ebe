#!/bin/bash
mkdir -p gpurun_out
K=${1:-conv1s_fwd}; ONLY=${2:-conv1_twin}; OUT=${3:-prof_x}
timeout 200 python tools/layer_bench.py --n 32768 --reps 1 --only $ONLY > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/$OUT python tools/layer_bench.py --n 32768 --reps 1 --only $ONLY > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
