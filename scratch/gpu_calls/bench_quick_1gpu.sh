#!/bin/bash
# quick check of bench.py itself after an edit: default line without the long legs, and the reference arm contract
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
tail -2 gpurun_out/quick_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick_bench.json').read())
print({k: d.get(k) for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','higher_is_better','scaling','vs_baseline','dtype','data','gpu_launches')})
print(d['roofline']); print(d['rooflines']['fused_full_sweep']); print(d['e2e']); print(d['clocks'])
PY
