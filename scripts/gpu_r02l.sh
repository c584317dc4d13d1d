#!/bin/bash
# multi-GPU: N-way PCIe probe + bench at N
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/pcie_probe_nway.py > gpurun_out/r02l_pcie_${N}way.json 2> gpurun_out/r02l_pcie_${N}way.err
cat gpurun_out/r02l_pcie_${N}way.json | head -40
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02l_bench_n$N.json 2> gpurun_out/r02l_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r02l_bench_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_n$N.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['e2e'])"
