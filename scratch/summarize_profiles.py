"""Turn the ncu outputs in gpurun_out/ into the small text summaries committed under profiles/."""
import collections, csv, subprocess, sys
out = open('profiles/launches_r1_summary.txt', 'w')
lines = [l for l in open('gpurun_out/launches_r1.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    name = r['Kernel Name'].split('(')[0][-70:]
    v = float(r['Metric Value'].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
out.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 1 (C2)\n")
out.write("# build() + passes of the step + kernel-timing passes; cold-cache serialized launch times: compare SHARES\n")
out.write("%10s %6s %7s  kernel\n" % ("total ms", "count", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.write("%10.3f %6d %6.1f%%  %s\n" % (v[1] / 1e6, v[0], 100 * v[1] / tot, k))
# one step = the launches from the last SpMM E = P.V (spmm_csr_kernel<2>) up to and including the merge that follows it
names = [r['Kernel Name'].split('(')[0][-70:] for r in rows]
vals = [float(r['Metric Value'].replace(',', '')) for r in rows]
starts = [i for i, n in enumerate(names) if 'spmm_csr_kernel<2>' in n and any('score_topk_tc' in m for m in names[i:i + 40])]
out.write("\n# one step (SpMM + fused scoring + merge), ms per kernel family\n")
if starts:
    i0 = starts[-1]
    i1 = next(i for i in range(i0, len(names)) if 'score_topk_tc' in names[i])
    while i1 + 1 < len(names) and 'merge' in names[i1 + 1]:
        i1 += 1
    fam = collections.OrderedDict()
    for i in range(i0, i1 + 1):
        key = 'DeviceRadixSort (CUB)' if ('identity_decomposer' in names[i] or 'Policy1000' in names[i]) else names[i].replace('<unnamed>::', '').replace('void ', '')
        fam[key] = fam.get(key, 0.0) + vals[i] / 1e6
    tot_step = sum(fam.values())
    for kname, v in fam.items():
        out.write("%10.3f ms %5.1f%%  %s\n" % (v, 100 * v / tot_step, kname))
    out.write("%10.3f ms total (%d launches)\n" % (tot_step, i1 - i0 + 1))
out.close()
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active', 'sm__pipe_tensor_subpipe_hmma', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster', 'lts__t_bytes.sum ',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.max', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum ', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'sm__inst_executed_pipe_tmem', 'smsp__inst_executed_pipe_uniform', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit', 'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
import os
for rep, dst in (('gpurun_out/prof_tc_r1.ncu-rep', 'profiles/score_topk_tc_r1_ncu.txt'), ('gpurun_out/prof_spmm_r1.ncu-rep', 'profiles/spmm_r1_ncu.txt'),
                 ('gpurun_out/prof_probe_r1.ncu-rep', 'profiles/probe_r1_ncu.txt')):
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rr[0], rr[1], rr[2]
    with open(dst, 'w') as f:
        f.write("# ncu --set full --clock-control none --import-source on (one launch, C2 workload) : %s\n" % rep)
        f.write("# kernel: %s\n" % vals[hdr.index('Kernel Name')][:160])
        for h, u, v in zip(hdr, units, vals):
            if any(w in h for w in want) and v != '':
                f.write("%-95s %-12s %s\n" % (h, u, v))
print(open('profiles/launches_r1_summary.txt').read())
