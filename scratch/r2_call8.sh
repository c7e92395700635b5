#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_c8_smi.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_models.py -q -x -k "shard or sampled or multi or row_sharded" > gpurun_out/r2_c8_tests.log 2>&1
tail -5 gpurun_out/r2_c8_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_c8_bench_n2.json 2> gpurun_out/r2_c8_bench_n2.err
tail -4 gpurun_out/r2_c8_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c8_bench_n2.json').read().strip().split('\n')[-1])
for k in ('value','ms_per_step','value_full_sweep','phase_ms','selfcheck','selfcheck_detail','build_s','e2e','sweep'):
    print('N2', k, d.get(k))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --scaling strong --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_c8_bench_n2_strong.json 2> gpurun_out/r2_c8_bench_n2_strong.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c8_bench_n2_strong.json').read().strip().split('\n')[-1])
for k in ('value','ms_per_step','phase_ms','selfcheck','scaling'):
    print('N2 strong', k, d.get(k))
PY
