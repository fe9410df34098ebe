#!/bin/bash
mkdir -p gpurun_out
# 1. launch list of one RND step at full size
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_rnd.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_rnd.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_rnd.log 2>&1
tail -n 1 gpurun_out/ncu_rnd.log | cut -c1-200
# 2. full capture of the dominant kernel (conv2 forward, window kernel) at n = 131072
timeout 200 python tools/layer_bench.py --reps 1 --only conv2_fwd > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_window_conv -s 2 -c 1 -f -o gpurun_out/prof_conv2_fwd_final python tools/layer_bench.py --reps 1 --only conv2_fwd > gpurun_out/ncu2.log 2>&1
tail -n 1 gpurun_out/ncu2.log
