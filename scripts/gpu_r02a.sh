#!/bin/bash
# round-2 check-in run: GPU tests, bench line, launch list of smoke() under ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02a_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke_ncu.log 2>&1
tail -3 gpurun_out/r02a_tests.log; head -c 1500 gpurun_out/r02a_bench.json; tail -2 gpurun_out/r02a_smoke_ncu.log
