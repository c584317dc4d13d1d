#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02r_tests.log
tail -4 gpurun_out/r02r_tests.log
out=gpurun_out/r02r_sweep.txt; : > $out
for k in -1 0 1 2; do for s in 1 0; do echo "== klass $k stage $s" >> $out; FB200_TRACE=1 FB200_INFLATE_STAGE=$s timeout 300 python scripts/prof_run.py 16384 3 $k >> $out 2>&1; done; done
grep -E "==|rep 2" $out | cut -c1-200
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
FB200_INFLATE_STAGE=1 timeout 600 ncu --metrics $M --clock-control none -k regex:k_inflate_par -c 1 --csv --log-file gpurun_out/r02r_m_stage1.csv python scripts/prof_run.py 16384 1 0 > gpurun_out/r02r_m_stage1.log 2>&1
FB200_INFLATE_STAGE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_inflate_par -c 1 -o gpurun_out/r02r_inflate_full python scripts/prof_run.py 16384 1 0 > gpurun_out/r02r_full.log 2>&1
