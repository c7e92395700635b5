#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29591 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c34_n8_weak.json 2> gpurun_out/r2_c34_n8_weak.err
timeout 400 $TR --master-port 29592 bench.py --gpus 8 --scaling strong --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c34_n8_strong.json 2> gpurun_out/r2_c34_n8_strong.err
python - <<'PY'
import json
for n in ('weak','strong'):
    try:
        d=json.loads(open(f'gpurun_out/r2_c34_n8_{n}.json').read().strip().split('\n')[-1])
        print(n, {k: d.get(k) for k in ('value','ms_per_step','phase_ms','selfcheck','sweep')}, (d.get('e2e') or {}).get('value'))
    except Exception as e: print(n, 'parse failed', e)
PY
