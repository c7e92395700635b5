#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_prof_all.txt
PB200_TC_PROF=gpurun_out/tc_prof_all.txt timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_c27.json 2> gpurun_out/r2_c27.err
tail -2 gpurun_out/r2_c27.err | cut -c1-200
python - <<'PY'
import numpy as np
blocks=[];cur=[]
for l in open('gpurun_out/tc_prof_all.txt'):
    if l.startswith('#'):
        if cur: blocks.append(cur)
        cur=[]
    else: cur.append([int(x) for x in l.split()])
if cur: blocks.append(cur)
labs=['cta','total','flush','wait_tfull','wait_afull','items','tiles','setup','max_flush','body','first_waits','survivors']
for bi,b in enumerate(blocks):
    a=np.array(b,dtype=np.int64)
    print('call',bi,' '.join('%s %.0f/%d' % (labs[j],a[:,j].mean(),a[:,j].max()) for j in (1,2,3,5,6,8,9,11)))
PY
