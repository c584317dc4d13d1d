#!/bin/bash
# per-access L2 policies on the parse tables / candidate bytes: timing sweep + DRAM bytes of each variant
mkdir -p gpurun_out
out=gpurun_out/r02m_sweep.txt; : > $out
V=$PWD/moonbit_flate_b200/variants
run() { echo "== $*" >> $out; env "$@" python scripts/prof_run.py 16384 3 >> $out 2>&1; }
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=25
for v in l2pol1 l2pol1h l2pol2 l2pol3; do
  run FB200_LIB=$V/libflate_b200_$v.so
  run FB200_LIB=$V/libflate_b200_$v.so FB200_PARSE_GWARPS=20
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for v in l2pol1 l2pol3; do
  FB200_LIB=$V/libflate_b200_$v.so timeout 600 ncu --metrics $M --clock-control none -k regex:k_parse -c 1 --csv --log-file gpurun_out/r02m_dram_$v.csv python scripts/prof_run.py 16384 1 > gpurun_out/r02m_dram_$v.log 2>&1
done
grep -E "==|rep 2" $out
