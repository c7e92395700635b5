set -x
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_b1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:probe_kernel -s 1 -c 1 -f -o gpurun_out/prof_probe_r1 $B > gpurun_out/ncu_b2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_topk_tc -s 1 -c 1 -f -o gpurun_out/prof_tc_r1 $B > gpurun_out/ncu_b3.log 2>&1
ls -la gpurun_out | tail
