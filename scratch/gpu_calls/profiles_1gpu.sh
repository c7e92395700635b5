#!/bin/bash
# final single-GPU verification of the tree: whole GPU suite, smoke, default bench (with e2e and the reference CPU arm),
# launch list and ncu captures of the kernels DESIGN.md quotes
mkdir -p gpurun_out
python -c "from polara_b200 import _build; print(_build.build())"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_c32_tests.log 2>&1
tail -3 gpurun_out/r2_c32_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_c32_bench.json 2> gpurun_out/r2_c32_bench.err
tail -2 gpurun_out/r2_c32_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c32_bench.json').read())
for k in ('value','ms_per_step','value_full_sweep','value_flat_norms','phase_ms','build_s','e2e','e2e_csr_fastpath','gpu_launches','clocks'):
    print(k, d.get(k))
print(d['roofline']); print(d['rooflines']['fused_full_sweep']); print({k: d['cpu_baseline'][k] for k in ('value','kind','cores')}, d['cpu_baseline'].get('settings'))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_c32_ref.json 2> gpurun_out/r2_c32_ref.err
cut -c1-600 gpurun_out/r2_c32_ref.json
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/r2_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:spmm_window" -s 0 -c 3 -f -o gpurun_out/prof_spmm_r2 $B > gpurun_out/r2_ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:spmm_window4" -c 1 -f -o gpurun_out/prof_spmm_step_r2 $B > gpurun_out/r2_ncu_e.log 2>&1
PB200_PRUNE=0 timeout 400 ncu --set full --clock-control none --import-source on -k "regex:score_topk_tc" -s 1 -c 1 -f -o gpurun_out/prof_tc_r2 $B > gpurun_out/r2_ncu_c.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:score_topk_tc|probe_kernel" -s 2 -c 2 -f -o gpurun_out/prof_tc_pruned_r2 $B > gpurun_out/r2_ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
