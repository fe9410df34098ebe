#!/bin/bash
# ncu --set full captures of the new policy-path kernels (one launch each) after the plain run exited 0
mkdir -p gpurun_out
timeout 300 python tools/policy_breakdown.py > gpurun_out/plain_policy.log 2>&1 || exit 1
cap() {  # kernel-regex  skip  output-name
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$3 python tools/policy_breakdown.py > gpurun_out/ncu_$3.log 2>&1
  tail -n 1 gpurun_out/ncu_$3.log | cut -c1-200
}
cap policy_heads_bwd 1 prof_policy_heads_bwd
cap policy_heads_fwd 1 prof_policy_heads_fwd
cap lstm_cell_fwd 20 prof_lstm_cell_fwd
cap lstm_cell_bwd 20 prof_lstm_cell_bwd
cap umma_gemm_kernel 12 prof_lstm_recurrent_gemm
ls -la gpurun_out/*.ncu-rep
