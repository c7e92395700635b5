#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "spmm or coo or rsvd or transpose" > gpurun_out/r2_c2_tests.log 2>&1
tail -15 gpurun_out/r2_c2_tests.log
timeout 600 python scratch/spmm_bench.py > gpurun_out/r2_c2_spmm_bench.txt 2>&1
cat gpurun_out/r2_c2_spmm_bench.txt | tail -45
