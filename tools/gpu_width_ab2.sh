#!/bin/bash
# width_kernel at 3.1 Gb: fewer resident blocks per SM (the grid caps them)
mkdir -p gpurun_out
for G in 5 4 3 2; do
  HSA_B200_WIDTH_GRID=$G timeout 600 python bench.py --reads-total 12500000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/wab.json 2> gpurun_out/wab.err || tail -3 gpurun_out/wab.err
  python - "$G" <<'PY'
import json,sys
j=json.load(open('gpurun_out/wab.json'))
lm={}
for nm,t in j['roofline']['launch_ms']: lm[nm]=lm.get(nm,0)+t
print("3.1 Gb width grid %s blocks/SM: %.3f M reads/s  ms/step %.1f  launches %s" % (sys.argv[1], j['value']/1e6, j['ms_per_step'], {k: round(v,1) for k,v in lm.items()}))
PY
done
