#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29595 bench.py --gpus 8 --users 10000000 --items 125000 --nnz 1000000000 --rank 128 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c35_bench_c3.json 2> gpurun_out/r2_c35_bench_c3.err
tail -3 gpurun_out/r2_c35_bench_c3.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_c35_bench_c3.json').read().strip().split('\n')[-1])
    for k in ('value','ms_per_step','value_full_sweep','phase_ms','selfcheck','build_s','sweep'):
        print('C3', k, d.get(k))
    print(d['rooflines']['fused_full_sweep']['frac'], d['rooflines']['spmm']['frac'])
except Exception as e: print('C3 parse failed', e)
PY
