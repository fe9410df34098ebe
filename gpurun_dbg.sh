#!/bin/bash
for d in 0 2 1 3; do echo "CB_CONV1_DEBUG=$d"; CB_CONV1_DEBUG=$d timeout 300 python tools/layer_bench.py --only conv1 2>&1 | grep -v wgrad; done
