#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c3_tests.log 2>&1
tail -15 gpurun_out/r2_c3_tests.log
timeout 600 python scratch/spmm_bench.py > gpurun_out/r2_c3_spmm_bench.txt 2>&1
grep -v "max rel" gpurun_out/r2_c3_spmm_bench.txt | tail -30
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c3_bench.json 2> gpurun_out/r2_c3_bench.err
tail -3 gpurun_out/r2_c3_bench.err; cat gpurun_out/r2_c3_bench.json
