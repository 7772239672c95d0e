#!/bin/bash
# default workload on the 3.1 Gb genome (one 12.5 M-read batch): drain budget vs the cooperative stage's share of the batch
mkdir -p gpurun_out
for B in 1000 2000 4000 8000; do
  HSA_B200_DRAIN_BUDGET=$B timeout 600 python bench.py --reads-total 12500000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/drain3g.json 2> gpurun_out/drain3g.err || tail -3 gpurun_out/drain3g.err
  python - <<PY
import json
j=json.load(open('gpurun_out/drain3g.json'))
r=j['roofline']
lm={}
for nm,t in r['launch_ms']: lm[nm]=lm.get(nm,0)+t
print("drain=$B: %.3f M reads/s  ms/batch %.1f  heavy %d  launches %s" % (j['value']/1e6, j['ms_per_step'], j['heavy_searches_handed_to_cooperative_kernel'], {k: round(v,1) for k,v in lm.items()}))
PY
done
