#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/bench_sweep.py --reads 10000000 \
  "HSA_B200_MINB=5" "HSA_B200_FORCE_ROWS=1 HSA_B200_MINB=5" "HSA_B200_FORCE_ROWS=1 HSA_B200_MINB=6" "HSA_B200_MINB=6" \
  "HSA_B200_DRAIN_BUDGET=3000" "HSA_B200_COOP_CHUNKS=1024" \
  > gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
