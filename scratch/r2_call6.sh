#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_c6_tests.log 2>&1
tail -30 gpurun_out/r2_c6_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c6_bench.json 2> gpurun_out/r2_c6_bench.err
tail -3 gpurun_out/r2_c6_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c6_bench.json').read())
for k in ('value','ms_per_step','value_full_sweep','value_flat_norms','phase_ms','build_s','build_detail','e2e','e2e_csr_fastpath'):
    print(k, d.get(k))
PY
timeout 300 python bench.py --users 1000000 --items 125000 --rank 128 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_c6_bench_r128.json 2> gpurun_out/r2_c6_bench_r128.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c6_bench_r128.json').read())
for k in ('value','ms_per_step','value_full_sweep','ms_per_step_full_sweep','phase_ms','sweep','build_s'):
    print('r128', k, d.get(k))
print(d['rooflines']['fused_full_sweep'])
PY
