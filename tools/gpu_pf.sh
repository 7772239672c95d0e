#!/bin/bash
mkdir -p gpurun_out
python tools/bench_sweep.py --reads 10000000 "" "HSA_B200_LIB=/root/repo/hsa_b200/lib_pf1.so" "HSA_B200_LIB=/root/repo/hsa_b200/lib_pf2.so" > gpurun_out/sweep_pf.log 2>&1
EXP_GENOME=3100000003 python tools/bench_sweep.py --reads 10000000 "" "HSA_B200_LIB=/root/repo/hsa_b200/lib_pf1.so" "HSA_B200_LIB=/root/repo/hsa_b200/lib_pf2.so" >> gpurun_out/sweep_pf.log 2>&1
cut -c1-250 gpurun_out/sweep_pf.log
