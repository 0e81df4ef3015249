"""Summarise one ncu report: key metrics, stall breakdown, hottest SASS lines per stall reason.
usage: python tools/ncu_stalls.py raw.csv source.csv [stall_col ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); h, u, r = rows[0], rows[1], rows[2]
for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
          'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers']:
    if k in h: print("%-75s %s %s" % (k, r[h.index(k)], u[h.index(k)]))
st = [(float(r[i]), h[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for i in range(len(h)) if 'smsp__average_warps_issue_stalled' in h[i] and h[i].endswith('_per_issue_active.ratio')]
tot = sum(v for v, _ in st)
print("stalls (warps per issue-active cycle, total %.2f): " % tot + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for v, k in sorted(st, reverse=True)[:10]))
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2]))); h = rows[1]; data = rows[2:]
    for col in sys.argv[3:] or ['stall_long_sb']:
        c = h.index(col); tot = sum(int(x[c] or 0) for x in data)
        print("--", col, tot, "of", sum(int(x[4] or 0) for x in data), "samples")
        for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][c] or 0))[:12]):
            print("  %5d %7s  %s" % (i, data[i][c], data[i][1].strip()[:80]))
