export PB200_TC_PAIR=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "score" 2>&1 | tail -15
timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/pair.err | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('PAIR', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
tail -5 gpurun_out/pair.err
