#!/bin/bash
# A/B of an environment knob on short benches at both genome sizes: $1 = VAR, $2.. = values
VAR=$1; shift
mkdir -p gpurun_out
for G in 3100000003 46000003; do
 for V in "$@"; do
  R=12500000; [ $G = 46000003 ] && R=10000000
  env $VAR=$V timeout 600 python bench.py --genome $G --reads-total $R --batch $R --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
j=json.load(open('gpurun_out/ab.json'))
r=j['roofline']
print("genome $G $VAR=$V: %.3f M reads/s  ms/step %.1f  launch_ms %s" % (j['value']/1e6, j['ms_per_step'], r['launch_ms'][:6]))
PY
 done
done
