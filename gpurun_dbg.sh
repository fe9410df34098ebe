#!/bin/bash
for d in 0 16 2 18; do echo "CB_CONV1_DEBUG=$d"; CB_CONV1_DEBUG=$d timeout 300 python tools/layer_bench.py --only c4_fwd 2>&1 | tail -1; done
