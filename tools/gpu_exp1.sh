#!/bin/bash
mkdir -p gpurun_out
python tools/exp_tail.py > gpurun_out/exp_tail.log 2>&1
cat gpurun_out/exp_tail.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o gpurun_out/search_full_10M -f \
    python bench.py --reads 10000000 --steps 1 --warmup 1 --no-cpu-baseline --no-probe > gpurun_out/ncu_full_10M.log 2>&1
tail -2 gpurun_out/ncu_full_10M.log
