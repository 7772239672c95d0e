#!/bin/bash
# headline workload, one 12.5 M-read batch: more resident blocks per SM made possible by fewer score buckets per lane
# (HSA_B200_MINB = launch bound / blocks per SM, HSA_B200_NB_FAST = score buckets of the fast kernel: 2 bytes of shared memory each per lane)
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --reads-total 12500000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/occ3g.json 2> gpurun_out/occ3g.err || tail -3 gpurun_out/occ3g.err
  python - "$*" <<'PY'
import json,sys
j=json.load(open('gpurun_out/occ3g.json'))
r=j['roofline']
lm={}
for nm,t in r['launch_ms']: lm[nm]=lm.get(nm,0)+t
print("%s: %.3f M reads/s  ms/batch %.1f  heavy %d  launches %s" % (sys.argv[1], j['value']/1e6, j['ms_per_step'], j['heavy_searches_handed_to_cooperative_kernel'], {k: round(v,1) for k,v in lm.items()}))
PY
}
run HSA_B200_MINB=5 HSA_B200_NB_FAST=64
run HSA_B200_MINB=6 HSA_B200_NB_FAST=32
run HSA_B200_MINB=6 HSA_B200_NB_FAST=40
run HSA_B200_MINB=6 HSA_B200_NB_FAST=48
run HSA_B200_MINB=8 HSA_B200_NB_FAST=32
run HSA_B200_MINB=5 HSA_B200_NB_FAST=32
