#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_sa.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_sa.log
tail -15 gpurun_out/pytest_gpu_sa.log
timeout 900 python tools/bench_sa.py > gpurun_out/bench_sa.json 2> gpurun_out/bench_sa.err; echo "sa rc=$?"
cat gpurun_out/bench_sa.json; tail -3 gpurun_out/bench_sa.err
