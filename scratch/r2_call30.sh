#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k item_sharded > gpurun_out/r2_c30_tests.log 2>&1
grep -n "Error\|assert\|line [0-9]*, in" gpurun_out/r2_c30_tests.log | head -40 | cut -c1-300
tail -3 gpurun_out/r2_c30_tests.log | cut -c1-300
