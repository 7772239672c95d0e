#!/bin/bash
# PUSH-phase split: parity, then a sweep of the PUSH vote threshold, then per-phase cycle counters (-DHSA_PHASE_PROF build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_sweep.py --reads 10000000 "HSA_B200_PUSH_MIN=1" "HSA_B200_PUSH_MIN=4" "HSA_B200_PUSH_MIN=8" "HSA_B200_PUSH_MIN=10" "HSA_B200_PUSH_MIN=12" "HSA_B200_PUSH_MIN=14" "HSA_B200_PUSH_MIN=10 HSA_B200_POP_BIAS=-8" "HSA_B200_PUSH_MIN=10 HSA_B200_POP_BIAS=-16" "HSA_B200_PUSH_MIN=10 HSA_B200_MINB=6" > gpurun_out/sweep14.log 2>&1
cat gpurun_out/sweep14.log
cp hsa_b200/libhsa_b200.so /tmp/lib_keep.so
HSA_B200_NVCC_EXTRA=-DHSA_PHASE_PROF python -m hsa_b200.build --force > /dev/null 2>&1
HSA_B200_TRACE=1 timeout 600 python tools/bench_sweep.py --reads 10000000 "HSA_B200_PUSH_MIN=1" "HSA_B200_PUSH_MIN=10" > gpurun_out/prof14.log 2>&1
cp /tmp/lib_keep.so hsa_b200/libhsa_b200.so
cat gpurun_out/prof14.log
