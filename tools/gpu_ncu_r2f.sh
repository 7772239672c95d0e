#!/bin/bash
# end of round 2: launch list + one full capture of the pass-1 search kernel at HEAD (3.1 Gb genome, one 12.5 M-read batch)
TAG=${1:-r02f}
mkdir -p gpurun_out
B="python bench.py --reads-total 12500000 --steps 1 --warmup 1 --no-cpu-baseline --no-probe --no-secondary"
timeout 900 $B > gpurun_out/ncu_plain_$TAG.json 2> gpurun_out/ncu_plain_$TAG.err; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"search_kernel|width_kernel|coop_kernel|bin_|repack_kernel" -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}_3g.csv $B > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 1800 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o gpurun_out/search_full_${TAG}_3g -f \
    $B > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | tail -5
