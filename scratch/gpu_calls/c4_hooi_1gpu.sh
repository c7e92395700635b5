#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_c13_tests.log 2>&1
tail -6 gpurun_out/r2_c13_tests.log | cut -c1-300
timeout 600 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/r2_c13_bench_c4.json 2> gpurun_out/r2_c13_bench_c4.err
tail -2 gpurun_out/r2_c13_bench_c4.err | cut -c1-300; cat gpurun_out/r2_c13_bench_c4.json | cut -c1-1200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c4_r2.csv python bench.py --config c4 --steps 2 --warmup 1 > gpurun_out/r2_c13_ncu_c4.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches_c4_r2.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for r in rows:
    n=r['Kernel Name'].split('(')[0][-60:]; v=float(r['Metric Value'].replace(',',''))
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print('%9.2f ms %5d %5.1f%% %s' % (v[1]/1e6, v[0], 100*v[1]/tot, k))
PY
timeout 900 python bench.py --config c5 --steps 2 > gpurun_out/r2_c13_bench_c5.json 2> gpurun_out/r2_c13_bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c13_bench_c5.json').read())
print('C5 build', d['build_s'], d['build_detail'])
for x in d['rank_sweep']: print({k: (round(v,3) if isinstance(v,float) else v) for k,v in x.items()})
PY
