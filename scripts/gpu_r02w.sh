#!/bin/bash
mkdir -p gpurun_out
V=$PWD/moonbit_flate_b200/variants
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02w_tests.log
tail -4 gpurun_out/r02w_tests.log
out=gpurun_out/r02w_sweep.txt; : > $out
for k in -1 0 1 2 3; do echo "== default klass $k" >> $out; timeout 300 python scripts/prof_run.py 16384 3 $k >> $out 2>&1; done
for v in pf0 pf2; do for k in -1 0 2; do echo "== $v klass $k" >> $out; FB200_LIB=$V/libflate_b200_$v.so timeout 300 python scripts/prof_run.py 16384 3 $k >> $out 2>&1; done; done
grep -E "==|rep 2" $out | cut -c1-190
