#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_icm.py -x -q -k "grouped or disagreement or icm" 2>&1 | tail -15
