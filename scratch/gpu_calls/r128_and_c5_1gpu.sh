#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_c17_tests.log 2>&1
tail -4 gpurun_out/r2_c17_tests.log | cut -c1-300
timeout 300 python bench.py --users 1000000 --items 125000 --rank 128 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_c17_bench_r128.json 2> gpurun_out/r2_c17_bench_r128.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c17_bench_r128.json').read())
print('r128 value %.3e step %.2f ms' % (d['value'], d['ms_per_step']), 'full %.2f ms' % d['ms_per_step_full_sweep'], 'fused_full %.2f ms frac %.3f' % (d['rooflines']['fused_full_sweep']['kernel_ms'], d['rooflines']['fused_full_sweep']['frac']))
PY
timeout 900 python bench.py --config c5 --steps 2 > gpurun_out/r2_c17_bench_c5.json 2> gpurun_out/r2_c17_bench_c5.err
tail -2 gpurun_out/r2_c17_bench_c5.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c17_bench_c5.json').read())
print('C5 build', d['build_s'])
for x in d['rank_sweep']: print({k: (round(v,3) if isinstance(v,float) else v) for k,v in x.items()})
PY
