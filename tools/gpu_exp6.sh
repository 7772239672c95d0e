#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"
python -c "
import json; j=json.load(open('gpurun_out/bench6.json')); print('value', j['value']/1e6, 'e2e', j['e2e']['value']/1e6, 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'])"
tail -3 gpurun_out/bench6.err
