#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "score or merge or topk" > gpurun_out/r2_c7_tests.log 2>&1
tail -3 gpurun_out/r2_c7_tests.log
for pi in 256 128; do
  PB200_PROBE=$pi timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c7_bench_p$pi.json 2> gpurun_out/r2_c7_bench_p$pi.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c7_bench_p$pi.json').read())
print('probe $pi', 'value %.3e step %.2f ms' % (d['value'], d['ms_per_step']), 'full %.2f ms' % d['ms_per_step_full_sweep'], d['phase_ms'], 'share %.4f' % d['sweep']['executed_share'],
      'fused_full %.2f ms frac %.3f' % (d['rooflines']['fused_full_sweep']['kernel_ms'], d['rooflines']['fused_full_sweep']['frac']), 'fused_default %.2f ms' % d['rooflines']['fused_default']['kernel_ms'], 'flat %.1f ms' % d['flat_norms']['ms_per_step'])
PY
done
