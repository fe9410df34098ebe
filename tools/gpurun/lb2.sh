#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_rnd.py tests/test_gpu_icm.py -x -q 2>&1 | tail -2
timeout 300 python tools/layer_bench.py --only conv1 2>&1 | tail -3
timeout 300 python tools/layer_bench.py --only c4 2>&1 | tail -2
