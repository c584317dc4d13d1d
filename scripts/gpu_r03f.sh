#!/bin/bash
# BASELINE configs[2] at full size (1 M small streams) and configs[4] through bench.py --workload, one GPU
mkdir -p gpurun_out
for w in small random runs; do
  timeout 1200 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r03f_bench_$w.json 2> gpurun_out/r03f_bench_$w.err; echo "$w rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r03f_bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['e2e']['value'] if 'e2e' in d else None, d['config']['workload'][:60])"
done
