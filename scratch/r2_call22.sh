#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "score" > gpurun_out/r2_c22_tests.log 2>&1
tail -3 gpurun_out/r2_c22_tests.log | cut -c1-300
for v in "" "PB200_TC_NOORDER=1" "PB200_TC_LOOKUP_FIRST=1" "PB200_TC_LOOKUP_FIRST=2"; do
  echo "== $v"
  env $v timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c22_bench.json 2> gpurun_out/r2_c22_bench.err
  tail -2 gpurun_out/r2_c22_bench.err | cut -c1-200
  python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c22_bench.json').read())
print({k: d.get(k) for k in ('value','ms_per_step','value_full_sweep','value_flat_norms','phase_ms')}, d['flat_norms'])
PY
done
