#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_c10_tests.log 2>&1
tail -6 gpurun_out/r2_c10_tests.log | cut -c1-300
timeout 400 python scratch/spmm_bench.py > gpurun_out/r2_c10_spmm_bench.txt 2>&1
grep -v "max rel" gpurun_out/r2_c10_spmm_bench.txt | tail -16
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c10_bench.json 2> gpurun_out/r2_c10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c10_bench.json').read())
for k in ('value','ms_per_step','value_full_sweep','phase_ms','build_s','build_detail','e2e','e2e_csr_fastpath'):
    print(k, d.get(k))
print(d['rooflines']['spmm'])
PY
timeout 600 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/r2_c10_bench_c4.json 2> gpurun_out/r2_c10_bench_c4.err
tail -2 gpurun_out/r2_c10_bench_c4.err | cut -c1-300; cat gpurun_out/r2_c10_bench_c4.json | cut -c1-1500
timeout 1200 python bench.py --config c5 --steps 2 > gpurun_out/r2_c10_bench_c5.json 2> gpurun_out/r2_c10_bench_c5.err
tail -2 gpurun_out/r2_c10_bench_c5.err | cut -c1-300; cat gpurun_out/r2_c10_bench_c5.json | cut -c1-3000
