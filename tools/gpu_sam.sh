#!/bin/bash
# SAM-field stage on the GPU box: bench line, then (after it exited 0) the launch list and one full capture per kernel
mkdir -p gpurun_out
timeout 900 python tools/bench_sam.py > gpurun_out/sam_bench_r02.json 2> gpurun_out/sam_bench_r02.err; echo "bench_sam rc=$?"
cat gpurun_out/sam_bench_r02.json; tail -3 gpurun_out/sam_bench_r02.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sam_|sel_|DeviceScan" -c 60 --csv --log-file gpurun_out/sam_launches_r02.csv \
    python tools/bench_sam.py --n 1000000 --steps 1 --sample 0 > gpurun_out/sam_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sam_dp|sam_pos|sel_chain|sel_finish" -c 4 -o gpurun_out/sam_full_r02 -f \
    python tools/bench_sam.py --n 1000000 --steps 0 --sample 0 > gpurun_out/sam_ncu_full.log 2>&1
ls -la gpurun_out | grep sam_
