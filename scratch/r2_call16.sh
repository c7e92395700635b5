#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "score" > gpurun_out/r2_c16_tests.log 2>&1
tail -3 gpurun_out/r2_c16_tests.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c16_bench.json 2> gpurun_out/r2_c16_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c16_bench.json').read())
print('value %.3e step %.2f ms' % (d['value'], d['ms_per_step']), 'full %.2f ms' % d['ms_per_step_full_sweep'], d['phase_ms'],
      'fused_full %.2f ms frac %.3f' % (d['rooflines']['fused_full_sweep']['kernel_ms'], d['rooflines']['fused_full_sweep']['frac']), 'fused_default %.2f ms' % d['rooflines']['fused_default']['kernel_ms'], 'flat %.1f ms' % d['flat_norms']['ms_per_step'])
PY
timeout 300 python bench.py --users 1000000 --items 125000 --rank 128 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_c16_bench_r128.json 2> gpurun_out/r2_c16_bench_r128.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c16_bench_r128.json').read())
print('r128 value %.3e step %.2f ms' % (d['value'], d['ms_per_step']), 'full %.2f ms' % d['ms_per_step_full_sweep'], 'fused_full %.2f ms frac %.3f' % (d['rooflines']['fused_full_sweep']['kernel_ms'], d['rooflines']['fused_full_sweep']['frac']))
PY
timeout 600 python bench.py --config c5 --scale 0.4 --steps 2 > gpurun_out/r2_c16_bench_c5s.json 2> gpurun_out/r2_c16_bench_c5s.err
tail -2 gpurun_out/r2_c16_bench_c5s.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c16_bench_c5s.json').read())
print('C5 x0.4 build', d['build_s'])
for x in d['rank_sweep']: print({k: (round(v,3) if isinstance(v,float) else v) for k,v in x.items()})
PY
