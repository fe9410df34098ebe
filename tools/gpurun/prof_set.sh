#!/bin/bash
# full ncu captures of the largest kernels (one launch each, after a plain run exited 0)
mkdir -p gpurun_out
cap() {  # kernel-regex  layer_bench-filter  output-name
  timeout 200 python tools/layer_bench.py --reps 1 --only $2 > gpurun_out/plain_$3.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -f -o gpurun_out/$3 python tools/layer_bench.py --reps 1 --only $2 > gpurun_out/ncu_$3.log 2>&1
  tail -n 1 gpurun_out/ncu_$3.log
}
cap conv1s_fwd conv1_twin prof_final_conv1_twin
cap umma_window_conv conv2_dgrad prof_final_conv2_dgrad
cap umma_gemm_kernel fc_fwd prof_final_fc_fwd
cap conv1s_wgrad conv1_wgrad prof_final_conv1_wgrad
