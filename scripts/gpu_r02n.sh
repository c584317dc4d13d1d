#!/bin/bash
# source staging by bulk asynchronous copies (FB_SRC_STAGE=1): parity first, then timing
mkdir -p gpurun_out
V=$PWD/moonbit_flate_b200/variants
FB200_LIB=$V/libflate_b200_stage.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 2gib" > gpurun_out/r02n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02n_tests.log
tail -4 gpurun_out/r02n_tests.log
out=gpurun_out/r02n_sweep.txt; : > $out
run() { echo "== $*" >> $out; env "$@" timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1; }
run FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=25
run FB200_LIB=$V/libflate_b200_stage.so
run FB200_LIB=$V/libflate_b200_stage.so FB200_PARSE_WARPS=4 FB200_PARSE_GWARPS=26
run FB200_LIB=$V/libflate_b200_stage.so FB200_PARSE_WARPS=4 FB200_PARSE_GWARPS=28
run FB200_LIB=$V/libflate_b200_stage.so FB200_PARSE_WARPS=5 FB200_PARSE_GWARPS=27
run FB200_PARSE_WARPS=4 FB200_PARSE_GWARPS=26
grep -E "==|rep 2" $out
