#!/bin/bash
for d in 0 1; do echo "epi debug=$d"; CB_EPI_DEBUG=$d timeout 120 python tools/layer_bench.py --only conv 2>&1 | grep -E "conv2_fwd|conv3_fwd|conv3_dgrad|conv2_dgrad"; done
