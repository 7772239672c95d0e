#!/bin/bash
# bit-plane sector layout: parity, timing, full ncu capture of the pass-1 search kernel (source-level)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
HSA_B200_TRACE=1 timeout 900 python tools/bench_sweep.py --reads 10000000 "" > gpurun_out/sweep15.log 2>&1
cat gpurun_out/sweep15.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o gpurun_out/search_r01_v5 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-probe > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
