#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shim.py -m gpu -x -q > gpurun_out/pytest_shim.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_shim.log
tail -30 gpurun_out/pytest_shim.log
