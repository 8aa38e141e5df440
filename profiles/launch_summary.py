#!/usr/bin/env python3
"""Per-kernel totals from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).
usage: python profiles/launch_summary.py gpurun_out/launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); mu = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split('(')[0].replace('east::', '').replace('void ', '')
    v = float(r[mv].replace(',', ''))
    v = v / 1e6 if r[mu] == 'ns' else (v / 1e3 if r[mu] == 'us' else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('# ncu launch list (cold-cache, serialised: compare SHARES, not absolutes); total %.3f ms over %d launches' % (
    tot, sum(a[0] for a in agg.values())))
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-40s launches=%5d total_ms=%9.3f avg_us=%9.1f share=%5.1f%%' % (k, c, ms, 1e3 * ms / c, 100 * ms / tot))
