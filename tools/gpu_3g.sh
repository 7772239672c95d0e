#!/bin/bash
# 3.1 Gb genome (configs[2] batches) on one GPU, then the other configs' throughput
mkdir -p gpurun_out
timeout 1500 python bench.py --genome 3100000003 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_3g_v9.json 2> gpurun_out/bench_3g_v9.err; echo "3g rc=$?"
cat gpurun_out/bench_3g_v9.json; tail -3 gpurun_out/bench_3g_v9.err
timeout 1200 python tools/bench_configs.py > gpurun_out/configs_v9.jsonl 2> gpurun_out/configs_v9.err; echo "rc=$?"
cat gpurun_out/configs_v9.jsonl; tail -3 gpurun_out/configs_v9.err
