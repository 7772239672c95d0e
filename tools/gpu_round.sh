#!/bin/bash
# GPU-box helper: gpu tests, default bench (both arms), ncu launch list and one full capture of the search kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"search_kernel|width_kernel|repack_kernel" -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-probe > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o gpurun_out/search_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-probe > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
