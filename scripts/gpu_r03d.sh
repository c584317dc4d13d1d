#!/bin/bash
mkdir -p gpurun_out
V=$PWD/moonbit_flate_b200/variants
out=gpurun_out/r03d_sweep.txt; : > $out
echo "== default" >> $out; timeout 300 python scripts/prof_run.py 16384 4 >> $out 2>&1
for v in lc12 lc20 sb768 sb1536 bs2 bs4 pw16 pw40; do echo "== $v" >> $out; FB200_LIB=$V/libflate_b200_$v.so timeout 300 python scripts/prof_run.py 16384 4 >> $out 2>&1; done
echo "== default again" >> $out; timeout 300 python scripts/prof_run.py 16384 4 >> $out 2>&1
grep -E "==|rep [23]" $out | cut -c1-170
