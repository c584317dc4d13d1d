#!/bin/bash
# round-2 parse experiments: warp mix sweep (time), DRAM bytes per mix (ncu metrics), full capture of k_parse
mkdir -p gpurun_out
out=gpurun_out/r02b_sweep.txt; : > $out
run() { echo "== $*" >> $out; env "$@" python scripts/prof_run.py 16384 3 >> $out 2>&1; }
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=25
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=20
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=15
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=10
run FB200_PARSE_WARPS=6 FB200_PARSE_GWARPS=14
run FB200_PARSE_WARPS=6 FB200_PARSE_GWARPS=8
run FB200_PARSE_WARPS=4 FB200_PARSE_GWARPS=28
python scripts/prof_run.py 16384 1 > gpurun_out/r02b_plain.log 2>&1 || exit 1
for g in 25 15 10; do
  FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=$g timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:k_parse -c 1 --csv --log-file gpurun_out/r02b_dram_g$g.csv python scripts/prof_run.py 16384 1 > gpurun_out/r02b_dram_g$g.log 2>&1
done
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_parse -c 1 -o gpurun_out/r02b_parse_full python scripts/prof_run.py 16384 1 > gpurun_out/r02b_full.log 2>&1
cat $out | grep -E "==|rep 2"
ls -la gpurun_out/r02b*
