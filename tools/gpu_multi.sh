#!/bin/bash
# N-GPU bench as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err; python -c "
import json,sys
j=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('n_gpus',j['n_gpus'],'value',j['value']/1e6,'e2e',j['e2e']['value']/1e6,'ms',j['ms_per_step'],'index_secs',j['index_build_secs'])"
