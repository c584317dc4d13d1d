#!/bin/bash
# multi-GPU bench with the pipelined frame exchange
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r03b_bench_n$N.json 2> gpurun_out/r03b_bench_n$N.err
echo "bench rc=$?"; tail -5 gpurun_out/r03b_bench_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r03b_bench_n$N.json')); print(d['value'], d['ms_per_step'], d.get('frame_exchange_ms_per_step'), d['roofline']['stage_ms_per_step'], d['e2e']['value'] if 'e2e' in d else None)"
