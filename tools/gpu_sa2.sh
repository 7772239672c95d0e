#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/bench_sa.py --genome 3100000003 --n 100000000 --cpu-sample 1000000 > gpurun_out/bench_sa_3g.json 2> gpurun_out/bench_sa_3g.err; echo "sa3g rc=$?"
cat gpurun_out/bench_sa_3g.json; tail -3 gpurun_out/bench_sa_3g.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_kernel -s 3 -c 1 -o gpurun_out/sa_full_r01 -f \
    python tools/bench_sa.py --cpu-sample 0 --steps 2 > gpurun_out/ncu_sa.log 2>&1
tail -2 gpurun_out/ncu_sa.log | cut -c1-200
