#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02q_sweep.txt; : > $out
for k in -1 0 2; do for s in 1 0; do echo "== klass $k stage $s" >> $out; FB200_TRACE=1 FB200_INFLATE_STAGE=$s timeout 300 python scripts/prof_run.py 16384 2 $k >> $out 2>&1; done; done
grep -E "==|rep 1|inflate:" $out | cut -c1-200
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
for s in 1 0; do
FB200_INFLATE_STAGE=$s timeout 600 ncu --metrics $M --clock-control none -k regex:k_inflate_par -c 1 --csv --log-file gpurun_out/r02q_m_stage$s.csv python scripts/prof_run.py 16384 1 0 > gpurun_out/r02q_m_stage$s.log 2>&1
done
FB200_INFLATE_STAGE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_inflate_par -c 1 -o gpurun_out/r02q_inflate_full python scripts/prof_run.py 16384 1 0 > gpurun_out/r02q_full.log 2>&1
