#!/bin/bash
# K6 staging (counting passes keep what they decode): parity, then timing per corpus class, with the staging off for comparison
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02p_tests.log
tail -4 gpurun_out/r02p_tests.log
out=gpurun_out/r02p_sweep.txt; : > $out
run() { echo "== $*" >> $out; env "$@" timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1; }
run FB200_INFLATE_STAGE=1
run FB200_INFLATE_STAGE=0
run FB200_INFLATE_CTAS=5
run FB200_INFLATE_CTAS=8
for k in 0 1 2 3; do echo "== klass $k" >> $out; timeout 300 python scripts/prof_run.py 16384 2 $k >> $out 2>&1; done
grep -E "==|rep [12]" $out
