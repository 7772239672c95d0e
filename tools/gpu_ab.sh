#!/bin/bash
# A/B of two builds of the library on short benches: $1 = alternative library (HSA_B200_LIB), both genome sizes
ALT=$1
mkdir -p gpurun_out
for G in 3100000003 46000003; do
 for L in "" "$ALT"; do
  R=12500000; [ $G = 46000003 ] && R=10000000
  HSA_B200_LIB=$L timeout 600 python bench.py --genome $G --reads-total $R --batch $R --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
j=json.load(open('gpurun_out/ab.json'))
r=j['roofline']
print("genome $G lib=${L:-default}: %.3f M reads/s  ms/step %.1f  launch_ms %s" % (j['value']/1e6, j['ms_per_step'], r['launch_ms'][:4]))
PY
 done
done
