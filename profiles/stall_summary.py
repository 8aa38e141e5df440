#!/usr/bin/env python3
"""Per-launch stall-reason mix + key memory metrics from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
def col(name):
    return hdr.index(name) if name in hdr else None
extra = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
         'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
         'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
         'lts__t_sectors.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[hdr.index('Kernel Name')].split('(')[0].replace('east::', '')
    st = [(h.replace('smsp__pcsamp_warps_issue_stalled_', ''), float(v)) for h, v in zip(hdr, r)
          if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and v not in ('', 'n/a')]
    tot = sum(v for _, v in st) or 1
    top = ' '.join('%s=%.0f%%' % (h, 100 * v / tot) for h, v in sorted(st, key=lambda x: -x[1])[:6])
    ex = ' '.join('%s=%s' % (e.split('.')[0].replace('smsp__', '').replace('l1tex__', 'l1_').replace('lts__', 'l2_'), r[col(e)]) for e in extra if col(e) is not None)
    print(name, '|', ex, '|', top)
