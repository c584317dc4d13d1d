#!/bin/bash
# helper warps for the parse (experiment): parity with helpers on, then warp-mix sweep
mkdir -p gpurun_out
FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=10 FB200_PARSE_HELPERS=15 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 2gib" > gpurun_out/r03a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r03a_tests.log
tail -3 gpurun_out/r03a_tests.log
out=gpurun_out/r03a_sweep.txt; : > $out
run() { echo "== $*" >> $out; env "$@" timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1; }
run FB200_PARSE_HELPERS=0
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=10 FB200_PARSE_HELPERS=0
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=10 FB200_PARSE_HELPERS=15
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=10 FB200_PARSE_HELPERS=8
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=15 FB200_PARSE_HELPERS=10
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=17 FB200_PARSE_HELPERS=10
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=20 FB200_PARSE_HELPERS=7
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=5 FB200_PARSE_HELPERS=10
grep -E "==|rep 2" $out | cut -c1-150
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=10 FB200_PARSE_HELPERS=15 timeout 600 ncu --metrics $M --clock-control none -k regex:k_parse -c 1 --csv --log-file gpurun_out/r03a_dram_h15.csv python scripts/prof_run.py 16384 1 > gpurun_out/r03a_dram_h15.log 2>&1
grep "k_parse" gpurun_out/r03a_dram_h15.csv | cut -d, -f13- | head
