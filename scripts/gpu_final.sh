#!/bin/bash
# final evidence of a round: tests, bench (+ reference arm), smoke, ncu launch list, --set full captures of K1 and K6.
# usage: bash scripts/gpu_final.sh <prefix>      (files land in gpurun_out/<prefix>_*)
P=${1:-r02_i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${P}_tests.log
tail -3 gpurun_out/${P}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${P}_smoke.log 2>&1; tail -1 gpurun_out/${P}_smoke.log
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err; echo "reference rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${P}_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['e2e']['value'], d['e2e'].get('serial',{}).get('value'), d['e2e'].get('platform_bound_gbs'))
print({k:(v.get('deflate_gbs'),v.get('inflate_gbs')) if isinstance(v,dict) else v for k,v in d.get('configs',{}).items()})
r=json.load(open('gpurun_out/${P}_bench_reference.json')); print('reference', r['value'], r.get('cpu_baseline'))
PY
# launch list (the same command exited 0 just above, without ncu)
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/${P}_bench_short.json 2> gpurun_out/${P}_bench_short.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${P}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# full captures of the two dominant kernels
timeout 300 python scripts/prof_run.py 16384 1 > gpurun_out/${P}_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_parse|k_inflate_par" -c 3 -o gpurun_out/${P}_full python scripts/prof_run.py 16384 1 > gpurun_out/${P}_full.log 2>&1
echo "full capture rc=$?"
python scripts/single_stream_probe.py > gpurun_out/${P}_single_stream.json 2> gpurun_out/${P}_single_stream.err; grep -E "mib|reader_ms|writer_ms" gpurun_out/${P}_single_stream.json
