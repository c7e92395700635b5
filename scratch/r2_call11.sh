#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2_c11_tests.log 2>&1
tail -8 gpurun_out/r2_c11_tests.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --users 2000000 --items 250000 --nnz 200000000 --rank 128 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2_c11_bench_c3like.json 2> gpurun_out/r2_c11_bench_c3like.err
tail -3 gpurun_out/r2_c11_bench_c3like.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_c11_bench_c3like.json').read().strip().split('\n')[-1])
    for k in ('value','ms_per_step','phase_ms','selfcheck','selfcheck_detail','build_s','sweep'):
        print('C3-like N2', k, d.get(k))
    print(d['rooflines']['spmm']['kernel'], d['rooflines']['spmm']['kernel_ms'])
except Exception as e: print('parse failed', e)
PY
