#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_models.py -q -k "operator or projector" > gpurun_out/verify_new.log 2>&1
tail -12 gpurun_out/verify_new.log | cut -c1-300
