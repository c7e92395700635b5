#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "score" > gpurun_out/r2_c28_tests.log 2>&1
tail -3 gpurun_out/r2_c28_tests.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c28_bench.json 2> gpurun_out/r2_c28_bench.err
tail -2 gpurun_out/r2_c28_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c28_bench.json').read())
print({k: d.get(k) for k in ('value','ms_per_step','value_full_sweep','value_flat_norms','phase_ms')}, d['flat_norms'])
print(d['rooflines']['fused_full_sweep']['kernel_ms'], d['rooflines']['fused_full_sweep']['frac'])
PY
