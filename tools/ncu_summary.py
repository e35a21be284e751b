"""tools/ncu_summary.py raw.csv [src.csv] -- key metrics (and per-source-line instruction counts) of an ncu capture."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__cycles_active.avg']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get('Kernel Name'))
    for k in keys:
        if k in d: print("  ", k, d[k], units[hdr.index(k)])
    st = [(float(d[k]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in hdr if 'issue_stalled' in k and 'per_issue_active' in k and d[k] not in ('', 'n/a')]
    print("   stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    tables = []; cur = None
    for r in rows:
        if r and r[0] == "File Path": cur = {'file': r[1], 'rows': []}; tables.append(cur)
        elif cur is not None: cur['rows'].append(r)
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    for t in tables:
        rr = t['rows']; h = None
        for i, r in enumerate(rr):
            if r and r[0] == "Line No": h = r; start = i + 1; break
        if not h: continue
        ii = h.index("Instructions Executed"); isamp = h.index("# Samples")
        tot = 0; lines = []
        for r in rr[start:]:
            if len(r) <= ii or not r[0]: continue
            try: v = int(r[ii])
            except ValueError: continue
            tot += v; lines.append((int(r[0]), v, r[isamp], r[1][:100]))
        print("=====", t['file'], "warp-instructions %.2fM" % (tot / 1e6))
        for l, v, sm, src in lines:
            if v / 1e6 >= thr: print("%8.2fM smp %6s L%d: %s" % (v / 1e6, sm, l, src))
