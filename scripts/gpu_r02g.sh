#!/bin/bash
mkdir -p gpurun_out
FB200_TRACE=2 timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu --no-configs > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_trace.txt
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench.json')); print(d['value'], d['e2e'])"
grep "fb200\] #" gpurun_out/r02g_trace.txt | tail -40
