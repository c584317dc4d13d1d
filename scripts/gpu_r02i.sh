#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02i_tests.log
tail -4 gpurun_out/r02i_tests.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02i_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step']); print(d['e2e']['value'], d['e2e']['ms_per_step'])
    for k,v in d['configs'].items(): print(k, v.get('deflate_ms'), v.get('inflate_ms'), v.get('deflate_stage_ms'), v.get('error'))
except Exception as e: print('bench failed', e)
PY
tail -3 gpurun_out/r02i_bench.err
