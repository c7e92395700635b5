"""Turn the round-2 ncu outputs in gpurun_out/ into the small text summaries committed under profiles/."""
import collections, csv, os, subprocess, sys
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r2"
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active',
        'sm__pipe_tensor_subpipe_hmma', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__cluster', 'lts__t_bytes.sum ', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__average_warps_issue_stalled',
        'sm__cycles_elapsed.max', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum ',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_tmem', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum ',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum ']


def launches(src, dst, header):
    lines = [l for l in open(src) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        name = r['Kernel Name'].split('(')[0][-70:]
        v = float(r['Metric Value'].replace(',', ''))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w') as out:
        out.write(header)
        out.write("%10s %6s %7s  kernel\n" % ("total ms", "count", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write("%10.3f %6d %6.1f%%  %s\n" % (v[1] / 1e6, v[0], 100 * v[1] / tot, k))
        names = [r['Kernel Name'].split('(')[0][-70:] for r in rows]
        vals = [float(r['Metric Value'].replace(',', '')) for r in rows]
        # one step = from the last SpMM launch (E = P V) through the merge after score_topk_tc
        starts = [i for i, n in enumerate(names) if 'spmm_window' in n and any('score_topk_tc' in m for m in names[i:i + 40])]
        out.write("\n# one step (SpMM + fused scoring + merge), ms per kernel family (cold-cache, serialised: compare shares)\n")
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


def report(rep, dst, note):
    if not os.path.exists(rep):
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    with open(dst, 'w') as f:
        f.write("# ncu --set full --clock-control none --import-source on : %s\n# %s\n" % (rep, note))
        for vals in rr[2:]:
            f.write("\n## kernel: %s   grid %s block %s\n" % (vals[hdr.index('Kernel Name')][:150], vals[hdr.index('Grid Size')] if 'Grid Size' in hdr else '', vals[hdr.index('Block Size')] if 'Block Size' in hdr else ''))
            for h, u, v in zip(hdr, units, vals):
                if any(w in h for w in want) and v != '':
                    f.write("%-95s %-12s %s\n" % (h, u, v))


if __name__ == "__main__":
    launches('gpurun_out/launches_%s.csv' % ROUND, 'profiles/launches_%s_summary.txt' % ROUND,
             "# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants (C2)\n"
             "# build() from host triplets + step passes; cold-cache serialised launch times: compare SHARES\n")
    report('gpurun_out/prof_spmm_%s.ncu-rep' % ROUND, 'profiles/spmm_%s_ncu.txt' % ROUND,
           "first three spmm_window_kernel<3> launches of build(): A.Q (ell 96, X = Q 38 MB), then panels 0 and 1 of A^T.W (X = W 384 MB in 10 panels)")
    report('gpurun_out/prof_tc_%s.ncu-rep' % ROUND, 'profiles/score_topk_tc_%s_ncu.txt' % ROUND,
           "PB200_PRUNE=0 (full sweep): score_topk_tc_kernel of the second step")
    report('gpurun_out/prof_tc_pruned_%s.ncu-rep' % ROUND, 'profiles/score_topk_tc_pruned_%s_ncu.txt' % ROUND,
           "default (sweep cut by the norm bound): probe_kernel and score_topk_tc_kernel of a step")
    report('gpurun_out/prof_spmm_step_%s.ncu-rep' % ROUND, 'profiles/spmm_step_%s_ncu.txt' % ROUND,
           "the step's SpMM E = P V (ell 64): spmm_window4_kernel<false> -- 128-bit gathers, half a warp per nnz")
    print(open('profiles/launches_%s_summary.txt' % ROUND).read())
