#!/bin/bash
# What a gpurun call of this round typically runs (tests, smoke, both bench arms); outputs under gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/run_gpu_checks.sh'                      (1 GPU)
#   gpurun --gpus N --timeout 1500 -- 'bash tools/run_gpu_checks.sh N'          (bench only, N ranks)
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
  timeout 300 python __graft_entry__.py smoke 2>&1 | grep smoke
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
  timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 400 gpurun_out/bench_1gpu.json
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
  tail -c 400 gpurun_out/bench_${N}gpu.json
fi
