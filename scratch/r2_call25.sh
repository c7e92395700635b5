#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:probe_kernel|head_bitmap_kernel" -s 2 -c 2 -f -o gpurun_out/prof_probe_r2 $B > gpurun_out/r2_ncu_p.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:score_topk_tc" -s 1 -c 1 -f -o gpurun_out/prof_tc_pruned_r2 $B > gpurun_out/r2_ncu_q.log 2>&1
ls -la gpurun_out/*.ncu-rep
