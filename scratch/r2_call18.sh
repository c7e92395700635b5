#!/bin/bash
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "score" > gpurun_out/r2_c18_tests.log 2>&1
tail -3 gpurun_out/r2_c18_tests.log | cut -c1-300
timeout 600 python bench.py --config c5 --scale 0.4 --steps 2 > gpurun_out/r2_c18_bench_c5s.json 2> gpurun_out/r2_c18_bench_c5s.err
tail -2 gpurun_out/r2_c18_bench_c5s.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c18_bench_c5s.json').read())
print('C5 x0.4 build', d['build_s'])
for x in d['rank_sweep']: print({k: (round(v,3) if isinstance(v,float) else v) for k,v in x.items()})
PY
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/r2_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:spmm_window" -s 0 -c 3 -f -o gpurun_out/prof_spmm_r2 $B > gpurun_out/r2_ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:spmm_window4_kernel<0>|spmm_window4_kernel<false>" -c 1 -f -o gpurun_out/prof_spmm_step_r2 $B > gpurun_out/r2_ncu_e.log 2>&1
PB200_PRUNE=0 timeout 400 ncu --set full --clock-control none --import-source on -k "regex:score_topk_tc|probe_kernel" -s 2 -c 2 -f -o gpurun_out/prof_tc_r2 $B > gpurun_out/r2_ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c18_bench.json 2> gpurun_out/r2_c18_bench.err
tail -2 gpurun_out/r2_c18_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c18_bench.json').read())
for k in ('value','ms_per_step','value_full_sweep','value_flat_norms','phase_ms','build_s','e2e','e2e_csr_fastpath'):
    print(k, d.get(k))
print(d['roofline']); print(d['rooflines']['fused_full_sweep']); print({k: d['cpu_baseline'][k] for k in ('value','kind','cores')}, d['cpu_baseline']['settings'])
PY
