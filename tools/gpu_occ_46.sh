#!/bin/bash
# the same knobs on the 46 Mb genome: configs[1] through bench.py, configs[0] / [3] / [4] through tools/bench_configs.py
mkdir -p gpurun_out
for S in "5 64" "6 40" "6 32"; do
  set -- $S
  export HSA_B200_MINB=$1 HSA_B200_NB_FAST=$2
  timeout 600 python bench.py --genome 46000003 --reads-total 10000000 --batch 10000000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-probe > gpurun_out/occ46.json 2> gpurun_out/occ46.err || tail -3 gpurun_out/occ46.err
  python - "$S" <<'PY'
import json,sys
j=json.load(open('gpurun_out/occ46.json'))
lm={}
for nm,t in j['roofline']['launch_ms']: lm[nm]=lm.get(nm,0)+t
print("46Mb minb,nb=%s: %.3f M reads/s  ms/step %.1f  heavy %d  launches %s" % (sys.argv[1], j['value']/1e6, j['ms_per_step'], j['heavy_searches_handed_to_cooperative_kernel'], {k: round(v,1) for k,v in lm.items()}))
PY
  timeout 600 python tools/bench_configs.py --genome 46000003 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('   ', str(j.get('config',''))[:90], '| kernel_ms', j.get('kernel_ms'), '| M reads/s kernel', round(j.get('reads_per_s_kernel',0)/1e6,2), '| heavy', j.get('heavy'))
"
done
