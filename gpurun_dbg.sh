#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_rnd.py -x -q 2>&1 | tail -1
for e in 8 12; do echo "CB_CONV1_EW=$e"; CB_CONV1_EW=$e timeout 300 python tools/layer_bench.py --only conv1_twin 2>&1 | tail -1; done
CB_CONV1_EW=12 timeout 600 python -m pytest tests/test_gpu_rnd.py -x -q 2>&1 | tail -1
