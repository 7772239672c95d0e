#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
HSA_B200_TRACE=1 timeout 1500 python bench.py --genome 3100000003 --reads 10000000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_3g.json 2> gpurun_out/bench_3g.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_3g.err | cut -c1-1200
python -c "
import json; j=json.loads(open('gpurun_out/bench_3g.json').read().strip().splitlines()[-1]); print('value', j['value']/1e6, 'e2e', j['e2e']['value']/1e6, 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'peak', j['roofline']['peak'], 'index_secs', j['index_build_secs'], 'aligned', j['aligned_fraction'], 'same', j['device_vs_host_path_identical'], 'heavy', j['heavy_searches_handed_to_cooperative_kernel'])"
