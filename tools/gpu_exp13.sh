#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_reads or rank" > gpurun_out/pytest_long.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_long.log
tail -12 gpurun_out/pytest_long.log
timeout 1200 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "rc=$?"
cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
