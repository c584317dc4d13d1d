#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02j_tests.log
tail -30 gpurun_out/r02j_tests.log
