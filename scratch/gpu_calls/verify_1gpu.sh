#!/bin/bash
# final check of a tree on one B200: the newest tests first (fast feedback), then the whole GPU suite and smoke()
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_models.py -q -x -k "operator or projector" > gpurun_out/verify_new.log 2>&1
tail -15 gpurun_out/verify_new.log | cut -c1-300
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/verify_tests.log 2>&1
tail -3 gpurun_out/verify_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
