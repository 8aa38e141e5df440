#!/usr/bin/env python3
"""One line per profiled launch from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'rdMB'),
        ('dram__bytes_write.sum', 'wrMB'), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'), ('launch__registers_per_thread', 'regs'),
        ('lts__t_sector_hit_rate.pct', 'L2hit%'), ('l1tex__t_sector_hit_rate.pct', 'L1hit%'),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'thr/inst'),
        ('sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed', 'xu%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'), ('launch__grid_size', 'grid')]
idx = [(hdr.index(w), n) for w, n in want if w in hdr]
units = rows[1]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    out = []
    for i, n in idx:
        v = r[i]
        if n == 'kernel':
            v = v.split('(')[0].replace('east::', '')
        elif n in ('rdMB', 'wrMB') and units[i] != 'Mbyte':
            v = v + units[i]
        elif n == 'time':
            v = v + units[i]
        out.append('%s=%s' % (n, v))
    print(' '.join(out))
