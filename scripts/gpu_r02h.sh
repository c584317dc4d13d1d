#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/r02h_bench_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r02h_bench_n$N.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['e2e'], d['config'].get('frame'))"
timeout 300 python -m pytest tests -m gpu -x -q -k "two_contexts or peer_frame" 2>&1 | tail -3
