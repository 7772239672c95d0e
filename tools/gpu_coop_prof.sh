#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coop_kernel -s 2 -c 1 -o gpurun_out/coop_full_r01 -f \
    python tools/exp_stress.py > gpurun_out/ncu_coop.log 2>&1
tail -3 gpurun_out/ncu_coop.log | cut -c1-200
