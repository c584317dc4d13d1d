#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02f_tests.log
tail -5 gpurun_out/r02f_tests.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02f_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step']); print(d.get('e2e')); print(json.dumps(d.get('configs'))[:1500]); print(d.get('cpu_baseline'))
except Exception as e: print('bench failed', e)
PY
tail -5 gpurun_out/r02f_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; head -c 600 gpurun_out/r02f_bench_ref.json
