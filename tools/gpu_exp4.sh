#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
HSA_B200_TRACE=1 timeout 600 python tools/exp_tail.py 2000000 10000000 > gpurun_out/exp_tail.log 2>&1
cat gpurun_out/exp_tail.log
timeout 1500 python tools/bench_sweep.py --reads 10000000 \
  "HSA_B200_MINB=5" "HSA_B200_MINB=6" "HSA_B200_MINB=4" \
  "HSA_B200_POP_BIAS=-4" "HSA_B200_POP_BIAS=-20" "HSA_B200_SLOW_MIN=3" "HSA_B200_SLOW_MIN=10" \
  "HSA_B200_MINB=6 HSA_B200_SLOW_MIN=3" \
  > gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
