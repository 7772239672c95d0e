#!/bin/bash
# N-GPU bench on the 3.1 Gb genome (configs[2] batches), launched as the driver launches bench.py
N=${1:-8}
mkdir -p gpurun_out
timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --genome 3100000003 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_3g_n$N.json 2> gpurun_out/bench_3g_n$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_3g_n$N.err; python -c "
import json,sys
j=json.loads(open('gpurun_out/bench_3g_n$N.json').read().strip().splitlines()[-1]); print('n_gpus',j['n_gpus'],'value',j['value']/1e6,'e2e',j['e2e']['value']/1e6,'ms',j['ms_per_step'],'index_secs',j['index_build_secs'])"
