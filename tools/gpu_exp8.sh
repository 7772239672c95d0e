#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
HSA_B200_TRACE=1 timeout 600 python tools/exp_tail.py 100000 2000000 10000000 > gpurun_out/exp_tail.log 2>&1
cat gpurun_out/exp_tail.log | cut -c1-900
