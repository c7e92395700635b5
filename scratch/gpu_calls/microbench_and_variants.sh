#!/bin/bash
# round-2 first GPU call: the two planned microbenchmarks + kernel variants already in the tree
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_c1_smi.txt
( timeout 120 ./scratch/mma_latency_bench ) > gpurun_out/r2_mma_latency.txt 2>&1
( timeout 120 ./scratch/pipe_skeleton ) > gpurun_out/r2_pipe_skeleton.txt 2>&1
for v in base allw pair; do
  unset PB200_TC_READOUT PB200_TC_PAIR
  [ $v = allw ] && export PB200_TC_READOUT=all
  [ $v = pair ] && export PB200_TC_PAIR=1
  timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "score" > gpurun_out/r2_c1_test_$v.log 2>&1
  timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c1_bench_$v.json 2> gpurun_out/r2_c1_bench_$v.err
done
unset PB200_TC_READOUT PB200_TC_PAIR
tail -3 gpurun_out/r2_c1_test_*.log
cat gpurun_out/r2_c1_bench_*.json | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])
    except Exception as e: print('bad', e)
"
