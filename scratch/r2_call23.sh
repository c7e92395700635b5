#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_prof_*.txt
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
PB200_TC_PROF=gpurun_out/tc_prof_order.txt PB200_TC_LOOKUP_FIRST=1 timeout 300 $B > gpurun_out/r2_c23_a.json 2> gpurun_out/r2_c23_a.err
PB200_TC_PROF=gpurun_out/tc_prof_noorder.txt PB200_TC_LOOKUP_FIRST=1 PB200_TC_NOORDER=1 timeout 300 $B > gpurun_out/r2_c23_b.json 2> gpurun_out/r2_c23_b.err
tail -2 gpurun_out/r2_c23_a.err | cut -c1-200
python - <<'PY'
import numpy as np
for name in ('order','noorder'):
    rows=[l.split() for l in open(f'gpurun_out/tc_prof_{name}.txt') if not l.startswith('#')]
    a=np.array(rows[-148:],dtype=np.int64)   # the last call
    print(name, 'calls', len(rows)//148)
    for j,lab in enumerate(['cta','total','flush','wait_tfull','wait_afull','items','tiles','setup','max_flush','body','first_waits','survivors']):
        if j==0: continue
        c=a[:,j]; print('  %-11s min %9d  mean %9.0f  max %9d' % (lab,c.min(),c.mean(),c.max()))
    o=np.argsort(-a[:,1])[:5]; print('  slowest CTAs:', a[o].tolist())
PY
