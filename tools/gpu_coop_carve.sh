#!/bin/bash
# coop_kernel: 8 blocks x 25.0 KB of shared memory = 200 KB asks for the 228 KB carve-out and leaves 28 KB of L1 -- the size at which
# the random-sector rate halves (probe carve-out sweep).  Preferred carve-outs that leave more L1 (and fewer resident blocks):
mkdir -p gpurun_out
for C in -1 86 72 57; do
  HSA_B200_COOP_CARVEOUT=$C timeout 600 python bench.py --reads-total 12500000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/cc.json 2> gpurun_out/cc.err || tail -3 gpurun_out/cc.err
  python - "$C" <<'PY'
import json,sys
j=json.load(open('gpurun_out/cc.json'))
lm={}
for nm,t in j['roofline']['launch_ms']: lm[nm]=lm.get(nm,0)+t
print("3.1 Gb coop carve-out %s %%: %.3f M reads/s  ms/step %.1f  launches %s" % (sys.argv[1], j['value']/1e6, j['ms_per_step'], {k: round(v,1) for k,v in lm.items()}))
PY
done
