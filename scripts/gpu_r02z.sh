#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02z_sweep.txt; : > $out
for g in 0 32 64 128; do echo "== L2_FETCH $g" >> $out; if [ $g = 0 ]; then timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1; else FB200_L2_FETCH=$g timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1; fi; done
grep -E "==|rep 2|granul" $out | cut -c1-190
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
FB200_L2_FETCH=32 timeout 600 ncu --metrics $M --clock-control none -k regex:"k_parse|k_inflate_par" -c 3 --csv --log-file gpurun_out/r02z_dram_f32.csv python scripts/prof_run.py 16384 1 > gpurun_out/r02z_dram_f32.log 2>&1
grep -E "k_parse<0>|k_inflate_par" gpurun_out/r02z_dram_f32.csv | cut -d, -f5,13- | head
