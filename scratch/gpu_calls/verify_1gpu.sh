#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2_c36_tests.log 2>&1
tail -3 gpurun_out/r2_c36_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r2_c36_bench.json 2> gpurun_out/r2_c36_bench.err
tail -2 gpurun_out/r2_c36_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c36_bench.json').read())
for k in ('value','ms_per_step','value_full_sweep','value_flat_norms','phase_ms','build_s','gpu_launches','clocks'):
    print(k, d.get(k))
print(d['e2e']['s_per_step'], d['e2e_csr_fastpath']['s_per_step'], d['roofline']['frac'], d['rooflines']['fused_full_sweep']['frac'], d['cpu_baseline']['value'])
PY
