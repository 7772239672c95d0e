#!/bin/bash
# width_kernel: launch bound (resident warps) and grid (one wave of resident blocks vs the first version's 8 blocks per SM)
mkdir -p gpurun_out
run() {
  G=$1; R=$2; shift 2
  env "$@" timeout 600 python bench.py --genome $G --reads-total $R --batch $R --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/wab.json 2> gpurun_out/wab.err || tail -3 gpurun_out/wab.err
  python - "$G $*" <<'PY'
import json,sys
j=json.load(open('gpurun_out/wab.json'))
lm={}
for nm,t in j['roofline']['launch_ms']: lm[nm]=lm.get(nm,0)+t
print("%s: %.3f M reads/s  ms/step %.1f  launches %s" % (sys.argv[1], j['value']/1e6, j['ms_per_step'], {k: round(v,1) for k,v in lm.items()}))
PY
}
for GR in "3100000003 12500000" "46000003 10000000"; do
  run $GR HSA_B200_WIDTH_MINB=5 HSA_B200_WIDTH_GRID=8
  run $GR HSA_B200_WIDTH_MINB=5 HSA_B200_WIDTH_GRID=0
  run $GR HSA_B200_WIDTH_MINB=6 HSA_B200_WIDTH_GRID=0
done
