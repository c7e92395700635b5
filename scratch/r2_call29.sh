#!/bin/bash
# 2 GPUs: multi-GPU parity tests, C2 weak (100K items per GPU) and strong (100K items total) lines with the self-check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2_c29_tests.log 2>&1
tail -4 gpurun_out/r2_c29_tests.log | cut -c1-400
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c29_n2_weak.json 2> gpurun_out/r2_c29_n2_weak.err
tail -2 gpurun_out/r2_c29_n2_weak.err | cut -c1-300
timeout 600 $T --master-port 29562 bench.py --gpus 2 --scaling strong --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c29_n2_strong.json 2> gpurun_out/r2_c29_n2_strong.err
tail -2 gpurun_out/r2_c29_n2_strong.err | cut -c1-300
python - <<'PY'
import json
for n in ('weak','strong'):
    try:
        d=json.loads(open(f'gpurun_out/r2_c29_n2_{n}.json').read().strip().split('\n')[-1])
        print(n, {k: d.get(k) for k in ('value','ms_per_step','phase_ms','selfcheck','build_s')}, (d.get('e2e') or {}).get('value'))
    except Exception as e: print(n, 'parse failed', e)
PY
