#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r03e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r03e_tests.log; tail -3 gpurun_out/r03e_tests.log
echo "== before"; FB200_LIB=$PWD/moonbit_flate_b200/variants/libflate_b200_head.so python scripts/reader_random_probe.py 1 16 2>&1 | tail -2
echo "== after"; python scripts/reader_random_probe.py 1 16 2>&1 | tail -2
