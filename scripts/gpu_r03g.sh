#!/bin/bash
# BASELINE configs[3] as stated: the fixed 8 GiB corpus over N GPUs (--scaling strong)
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --scaling strong --steps 5 --warmup 3 > gpurun_out/r03g_bench_strong_n$N.json 2> gpurun_out/r03g_bench_strong_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r03g_bench_strong_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r03g_bench_strong_n$N.json')); print(d['value'], d['ms_per_step'], d.get('frame_exchange_ms_per_step'), d['scaling'], d['config']['workload'][:80], d['e2e']['value'] if 'e2e' in d else None)"
