#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02s_tests.log
tail -4 gpurun_out/r02s_tests.log
out=gpurun_out/r02s_sweep.txt; : > $out
for k in -1 0 1 2 3 4 5; do echo "== klass $k" >> $out; timeout 300 python scripts/prof_run.py 16384 3 $k >> $out 2>&1; done
echo "== CTAS=5" >> $out; FB200_INFLATE_CTAS=5 timeout 300 python scripts/prof_run.py 16384 3 >> $out 2>&1
grep -E "==|rep 2" $out | cut -c1-200
python scripts/single_stream_probe.py > gpurun_out/r02s_single.json 2> gpurun_out/r02s_single.err; grep -E "mib|reader_ms|writer_ms" gpurun_out/r02s_single.json
