#!/bin/bash
mkdir -p gpurun_out
V=$PWD/moonbit_flate_b200/variants
out=gpurun_out/r03i_sweep.txt; : > $out
echo "== default" >> $out; timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1
for v in dw40 dw80; do echo "== $v" >> $out; FB200_LIB=$V/libflate_b200_$v.so timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1; done
grep -E "==|rep 2" $out | cut -c1-120
