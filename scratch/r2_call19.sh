#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_c19_tests.log 2>&1
tail -4 gpurun_out/r2_c19_tests.log | cut -c1-300
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:spmm_window4" -c 1 -f -o gpurun_out/prof_spmm_step_r2 $B > gpurun_out/r2_ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
