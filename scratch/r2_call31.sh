#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29571 bench.py --gpus 8 --users 10000000 --items 125000 --nnz 1000000000 --rank 128 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c31_bench_c3.json 2> gpurun_out/r2_c31_bench_c3.err
tail -6 gpurun_out/r2_c31_bench_c3.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_c31_bench_c3.json').read().strip().split('\n')[-1])
    for k in ('value','ms_per_step','value_full_sweep','ms_per_step_full_sweep','phase_ms','selfcheck','build_s','build_detail','sweep','nnz_actual'):
        print('C3', k, d.get(k))
    print('C3 roofs', d['rooflines']['spmm'], d['rooflines']['fused_full_sweep'])
except Exception as e: print('C3 parse failed', e)
PY
timeout 600 $TR --master-port 29572 bench.py --gpus 8 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c31_bench_n8_strong.json 2> gpurun_out/r2_c31_bench_n8_strong.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_c31_bench_n8_strong.json').read().strip().split('\n')[-1])
    for k in ('value','ms_per_step','phase_ms','selfcheck','e2e'):
        print('N8 strong', k, d.get(k))
except Exception as e: print('strong parse failed', e)
PY
timeout 600 $TR --master-port 29573 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c31_bench_n8_weak.json 2> gpurun_out/r2_c31_bench_n8_weak.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_c31_bench_n8_weak.json').read().strip().split('\n')[-1])
    for k in ('value','ms_per_step','phase_ms','selfcheck','e2e'):
        print('N8 weak', k, d.get(k))
except Exception as e: print('weak parse failed', e)
PY
