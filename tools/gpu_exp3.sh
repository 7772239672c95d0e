#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o gpurun_out/search_v3_10M -f \
    python bench.py --reads 10000000 --steps 1 --warmup 1 --no-cpu-baseline --no-probe > gpurun_out/ncu_v3.log 2>&1
tail -2 gpurun_out/ncu_v3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:width_kernel -s 2 -c 1 -o gpurun_out/width_v3_10M -f \
    python bench.py --reads 10000000 --steps 1 --warmup 1 --no-cpu-baseline --no-probe > gpurun_out/ncu_v3w.log 2>&1
tail -2 gpurun_out/ncu_v3w.log
python bench.py --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; cat gpurun_out/bench_v3.json
