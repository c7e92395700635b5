#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "shard or probe or heavy" > gpurun_out/r2_c33_k.log 2>&1
tail -3 gpurun_out/r2_c33_k.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2_c33_tests.log 2>&1
tail -3 gpurun_out/r2_c33_tests.log | cut -c1-400
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29581 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c33_n2_weak.json 2> gpurun_out/r2_c33_n2_weak.err
tail -2 gpurun_out/r2_c33_n2_weak.err | cut -c1-300
timeout 600 $T --master-port 29582 bench.py --gpus 2 --scaling strong --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c33_n2_strong.json 2> gpurun_out/r2_c33_n2_strong.err
python - <<'PY'
import json
for n in ('weak','strong'):
    try:
        d=json.loads(open(f'gpurun_out/r2_c33_n2_{n}.json').read().strip().split('\n')[-1])
        print(n, {k: d.get(k) for k in ('value','ms_per_step','phase_ms','selfcheck','build_s')}, (d.get('e2e') or {}).get('value'))
    except Exception as e: print(n, 'parse failed', e)
PY
