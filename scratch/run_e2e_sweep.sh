set -x
python -m pytest tests/test_gpu_models.py -x -q 2>&1 | tail -2
for c in 2 4 6 8; do
  PB200_STREAM_CHUNKS=$c python bench.py --steps 5 --warmup 3 2>gpurun_out/e2e_$c.err | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('chunks',$c, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['s_per_step'])"
done
