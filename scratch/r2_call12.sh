#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "slab or large_rank or spmm_matches or early_termination" > gpurun_out/r2_c12_tests.log 2>&1
tail -4 gpurun_out/r2_c12_tests.log | cut -c1-300
timeout 1200 python bench.py --config c5 --steps 2 > gpurun_out/r2_c12_bench_c5.json 2> gpurun_out/r2_c12_bench_c5.err
tail -2 gpurun_out/r2_c12_bench_c5.err | cut -c1-300; cat gpurun_out/r2_c12_bench_c5.json | cut -c1-3500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_r2.csv python bench.py --config c4 --steps 2 --warmup 1 --scale 1.0 > gpurun_out/r2_c12_ncu_c4.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches_c4_r2.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for r in rows:
    n=r['Kernel Name'].split('(')[0][-60:]; v=float(r['Metric Value'].replace(',',''))
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print('%9.2f ms %4d %5.1f%% %s' % (v[1]/1e6, v[0], 100*v[1]/tot, k))
PY
