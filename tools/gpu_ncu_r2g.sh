#!/bin/bash
# end of round 2: full captures of the two other kernels of the step at the headline workload (3.1 Gb, one 12.5 M-read batch)
mkdir -p gpurun_out
B="python bench.py --reads-total 12500000 --steps 1 --warmup 1 --no-cpu-baseline --no-probe --no-secondary"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:width_kernel -s 10 -c 1 -o gpurun_out/width_full_r02g_3g -f \
    $B > gpurun_out/ncu_width_r02g.log 2>&1; echo "width rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:coop_kernel -s 8 -c 1 -o gpurun_out/coop_full_r02g_3g -f \
    $B > gpurun_out/ncu_coop_r02g.log 2>&1; echo "coop rc=$?"
ls -la gpurun_out/*r02g*
