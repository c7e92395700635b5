#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c5_tests.log 2>&1
tail -4 gpurun_out/r2_c5_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c5_bench.json 2> gpurun_out/r2_c5_bench.err
tail -3 gpurun_out/r2_c5_bench.err; cat gpurun_out/r2_c5_bench.json | head -c 6000
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/r2_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_window_kernel -s 0 -c 3 -f -o gpurun_out/prof_spmm_r2 $B > gpurun_out/r2_ncu_b.log 2>&1
PB200_PRUNE=0 timeout 400 ncu --set full --clock-control none --import-source on -k "regex:score_topk_tc|probe_kernel" -s 2 -c 2 -f -o gpurun_out/prof_tc_r2 $B > gpurun_out/r2_ncu_c.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:score_topk_tc" -s 2 -c 1 -f -o gpurun_out/prof_tc_pruned_r2 $B > gpurun_out/r2_ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
