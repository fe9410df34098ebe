#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_umma.py tests/test_gpu_rnd.py tests/test_gpu_icm.py -k "conv1 or golden or oracle" -x -q 2>&1 | tail -3
timeout 120 python tools/layer_bench.py --only conv1 2>&1 | tail -3
for d in 1 2 3; do echo -n "debug=$d "; CB_CONV1_DEBUG=$d timeout 120 python tools/layer_bench.py --only conv1_twin 2>&1 | tail -1; done
