#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
free -g | head -2
HSA_B200_TRACE=1 timeout 1500 python bench.py --genome 3100000003 --reads 10000000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_3g.json 2> gpurun_out/bench_3g.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_3g.err | cut -c1-700
python -c "
import json; j=json.loads(open('gpurun_out/bench_3g.json').read().strip().splitlines()[-1]); print('value', j['value']/1e6, 'e2e', j['e2e']['value']/1e6, 'ms', j['ms_per_step'], 'roofline', j['roofline'], 'index_secs', j['index_build_secs'], 'aligned', j['aligned_fraction'])"
