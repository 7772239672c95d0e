#!/bin/bash
# L2 fetch granularity experiment: probe + short 3.1 Gb bench at 32 / 64 / 128 / driver default
mkdir -p gpurun_out
for G in 32 64 128 0; do
  HSA_B200_L2_FETCH=$G timeout 600 python bench.py --reads-total 12500000 --steps 3 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/l2fetch_$G.json 2> gpurun_out/l2fetch_$G.err
  python - <<PY
import json
j=json.load(open('gpurun_out/l2fetch_$G.json'))
r=j['roofline']
print("fetch=$G value %.3f M reads/s  ms/step %.1f  probe %s  launch_ms %s" % (j['value']/1e6, j['ms_per_step'], {k: round(v) for k,v in r['random_sector_probe']['variants_gbs'].items()}, r['launch_ms'][:4]))
PY
done
