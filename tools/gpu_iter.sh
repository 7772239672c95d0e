#!/bin/bash
# one optimisation iteration on the GPU box: parity, timing of the 10 M-read batch (+ trace), optional full ncu capture
# usage: gpu_iter.sh <tag> [ncu] [variants...]
TAG=${1:-iter}; shift
NCU=0; if [ "$1" == "ncu" ]; then NCU=1; shift; fi
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
if [ $# -eq 0 ]; then set -- ""; fi
HSA_B200_TRACE=1 timeout 1200 python tools/bench_sweep.py --reads 10000000 "$@" > gpurun_out/sweep_$TAG.log 2>&1
cut -c1-900 gpurun_out/sweep_$TAG.log
if [ $NCU -eq 1 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o gpurun_out/search_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-probe > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
fi
