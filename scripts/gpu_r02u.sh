#!/bin/bash
mkdir -p gpurun_out
V=$PWD/moonbit_flate_b200/variants
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02u_tests.log
tail -4 gpurun_out/r02u_tests.log
out=gpurun_out/r02u_sweep.txt; : > $out
for k in -1 0 1 2 3 4 5; do echo "== default klass $k" >> $out; FB200_TRACE=1 timeout 300 python scripts/prof_run.py 16384 3 $k >> $out 2>&1; done
for v in big0 big9; do for k in -1 0 2 3; do echo "== $v klass $k" >> $out; FB200_LIB=$V/libflate_b200_$v.so timeout 300 python scripts/prof_run.py 16384 3 $k >> $out 2>&1; done; done
grep -E "==|rep 2|rounds" $out | cut -c1-190 | uniq
