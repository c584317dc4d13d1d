#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02v_sweep.txt; : > $out
for o in 1 0; do echo "== order $o" >> $out; FB200_INFLATE_ORDER=$o timeout 300 python scripts/prof_run.py 16384 4 >> $out 2>&1; done
grep -E "==|rep [23]" $out | cut -c1-190
python scripts/pcie_probe_nway.py > gpurun_out/r02v_pcie_1way.json 2>&1; tail -5 gpurun_out/r02v_pcie_1way.json
# stall profile of K6 on incompressible data
timeout 300 python scripts/prof_run.py 16384 1 2 > gpurun_out/r02v_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_inflate_par -c 1 -o gpurun_out/r02v_inflate_random_full python scripts/prof_run.py 16384 1 2 > gpurun_out/r02v_full.log 2>&1
