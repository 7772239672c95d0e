#!/bin/bash
# L2 persisting-window experiment on the 46 Mb (L2-resident) index: search kernel with and without the access policy window
mkdir -p gpurun_out
for V in 0 1 0 1; do
  env HSA_B200_L2_PERSIST=$V timeout 600 python bench.py --genome 46000003 --reads-total 10000000 --batch 10000000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - "$V" <<'PY'
import json,sys
j=json.load(open('gpurun_out/ab.json')); r=j['roofline']
print("46 Mb HSA_B200_L2_PERSIST=%s: %.3f M reads/s  ms/step %.1f  launch_ms %s" % (sys.argv[1], j['value']/1e6, j['ms_per_step'], r.get('launch_ms', [])[:5]), flush=True)
PY
done
