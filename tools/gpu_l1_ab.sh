#!/bin/bash
# L1-size experiment (round 2): search kernel with different shared-memory footprints / carve-outs, both genome sizes
mkdir -p gpurun_out
run() {
 for G in 46000003 3100000003; do
  R=12500000; [ $G = 46000003 ] && R=10000000
  env "$@" timeout 600 python bench.py --genome $G --reads-total $R --batch $R --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - "$G" "$*" <<'PY'
import json,sys
try:
    j=json.load(open('gpurun_out/ab.json')); r=j['roofline']
    print("genome %s [%s]: %.3f M reads/s  ms/step %.1f  launch_ms %s" % (sys.argv[1], sys.argv[2], j['value']/1e6, j['ms_per_step'], r.get('launch_ms', [])[:6]), flush=True)
except Exception as e:
    print("genome %s [%s]: FAILED %s" % (sys.argv[1], sys.argv[2], e), flush=True)
PY
 done
}
run HSA_X=0
run HSA_B200_FORCE_ROWS=1
run HSA_B200_CARVEOUT=100
run HSA_B200_FORCE_ROWS=1 HSA_B200_CARVEOUT=40
run HSA_B200_BLOCKS_PER_SM=4 HSA_B200_CARVEOUT=66
run HSA_B200_NB_FAST=48
run HSA_B200_NB_FAST=32 HSA_B200_CARVEOUT=64
