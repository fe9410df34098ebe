#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_umma.py 2>&1 | tee gpurun_out/probe.log
