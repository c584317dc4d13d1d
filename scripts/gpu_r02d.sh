#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02d_tests.log
tail -5 gpurun_out/r02d_tests.log
out=gpurun_out/r02d_sweep.txt; : > $out
run() { echo "== $*" >> $out; env "$@" timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1; }
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=23
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=20
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=15
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=10
run FB200_PARSE_WARPS=6 FB200_PARSE_GWARPS=12
run FB200_PARSE_WARPS=4 FB200_PARSE_GWARPS=24
run FB200_LIB=$PWD/moonbit_flate_b200/variants/libflate_b200_lb1024.so
run FB200_LIB=$PWD/moonbit_flate_b200/variants/libflate_b200_lb1024.so FB200_PARSE_GWARPS=20
grep -E "==|rep 2|rror" $out
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02d_bench.json')); print(d['value'], d['e2e'])"; tail -3 gpurun_out/r02d_bench.err
python scripts/prof_run.py 16384 1 > gpurun_out/r02d_plain.log 2>&1 || exit 1
for g in 23 10; do
  FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=$g timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:k_parse -c 1 --csv --log-file gpurun_out/r02d_dram_g$g.csv python scripts/prof_run.py 16384 1 > gpurun_out/r02d_dram_g$g.log 2>&1
done
