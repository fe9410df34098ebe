#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_wgrad1.py 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/probe_w1.log
timeout 300 python tools/layer_bench.py --only conv1 2>&1 | tail -3
