#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02c_sweep.txt; : > $out
run() { echo "== $*" >> $out; env "$@" python scripts/prof_run.py 16384 3 >> $out 2>&1; }
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=25
run FB200_LIB=$PWD/moonbit_flate_b200/variants/libflate_b200_lowprio.so
run FB200_PARSE_WARPS=6 FB200_PARSE_GWARPS=24
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=27
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02c_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
python scripts/prof_run.py 16384 1 > gpurun_out/r02c_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_inflate_par -c 1 -o gpurun_out/r02c_inflate_full python scripts/prof_run.py 16384 1 > gpurun_out/r02c_full.log 2>&1
grep -E "==|rep 2" $out; tail -3 gpurun_out/r02c_tests.log; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench.json')); print(d['value'], d['e2e'])"; tail -3 gpurun_out/r02c_bench.err
