#!/bin/bash
# round 2, first GPU call: the gpu tests (incl. bench-scale parity against the reference binary), then a short bench of
# both arms on the 3.1 Gb workload
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -25 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --reads-total 25000000 --steps 2 --warmup 1 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err
