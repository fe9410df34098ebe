#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_umma.py tests/test_gpu_rnd.py tests/test_gpu_icm.py -k "conv1 or golden or oracle" -x -q 2>&1 | tail -3
timeout 120 python tools/layer_bench.py --only conv1 2>&1 | tail -3
