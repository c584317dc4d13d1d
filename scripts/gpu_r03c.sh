#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "writer or multi_block or long or stream_bytes or 2gib" > gpurun_out/r03c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r03c_tests.log; tail -3 gpurun_out/r03c_tests.log
python scripts/writer_stage_probe.py 1 16 128 2>&1 | tail -4
python scripts/single_stream_probe.py 1 16 > gpurun_out/r03c_single.json 2>/dev/null; grep -E "mib|writer_ms|one_call" gpurun_out/r03c_single.json
