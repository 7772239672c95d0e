#!/bin/bash
mkdir -p gpurun_out
HSA_B200_LIB=$PWD/hsa_b200/libhsa_b200_prof.so.keep HSA_B200_TRACE=1 timeout 600 python tools/exp_tail.py 10000000 > gpurun_out/exp_prof.log 2>&1
tail -2 gpurun_out/exp_prof.log | cut -c1-1500
for db in 1000 4000 8000; do
HSA_B200_DRAIN_BUDGET=$db HSA_B200_TRACE=1 timeout 600 python tools/exp_tail.py 10000000 2>&1 | tail -2 | cut -c1-700
done
