#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02y_tests.log
tail -4 gpurun_out/r02y_tests.log
out=gpurun_out/r02y_sweep.txt; : > $out
for k in -1 0 1 2 3 4 5; do echo "== default klass $k" >> $out; timeout 300 python scripts/prof_run.py 16384 3 $k >> $out 2>&1; done
grep -E "==|rep 2" $out | cut -c1-190
