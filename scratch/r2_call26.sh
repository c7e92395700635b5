#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_c26_tests.log 2>&1
tail -4 gpurun_out/r2_c26_tests.log | cut -c1-400
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c26_bench.json 2> gpurun_out/r2_c26_bench.err
tail -2 gpurun_out/r2_c26_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c26_bench.json').read())
print({k: d.get(k) for k in ('value','ms_per_step','value_full_sweep','value_flat_norms','phase_ms')}, d['flat_norms'])
print(d['rooflines']['fused_full_sweep']['kernel_ms'], d['rooflines']['fused_full_sweep']['frac'])
PY
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/r2_ncu_a.log 2>&1
python scratch/summarize_profiles_r2.py r2 > /dev/null 2>&1; tail -24 profiles/launches_r2_summary.txt
