#!/bin/bash
# stress configuration (150 bp, n5 o2): how the drain budget (searches handed to the cooperative kernel once the work queue is dry)
# trades the fast kernel's serial tail against cooperative-kernel time; 500 k and 2 M reads
for N in 500000 2000000; do
 for B in 1000 2000 4000 8000 16000 0; do
  HSA_B200_DRAIN_BUDGET=$B EXP_N=$N timeout 300 python tools/exp_stress.py 2>/dev/null | grep stress | sed "s/^/drain=$B n=$N: /"
 done
done
