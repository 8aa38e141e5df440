#!/usr/bin/env python3
"""Print the hottest SASS instructions (by warp-stall samples) of the first kernel in an ncu report.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass | python profiles/sass_hot.py [N]"""
import csv, sys
n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 14
rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
hdr = rows[hi]; src = hdr.index('Source'); ex = hdr.index('Instructions Executed'); samp = hdr.index('Warp Stall Sampling (All Samples)')
def f(x):
    try: return int(float(x))
    except ValueError: return 0
body = []
for r in rows[hi + 1:]:
    if len(r) > samp:
        if r[0] == 'Kernel Name': break
        body.append(r)
tot = sum(f(r[samp]) for r in body) or 1
top = sorted(range(len(body)), key=lambda i: -f(body[i][samp]))[:n_top]
for i in sorted(top):
    print('---- line %d: %.1f%% of samples' % (i, 100.0 * f(body[i][samp]) / tot))
    for j in range(max(0, i - 3), i + 1):
        print('    %5d  %-60s samples=%s exec=%s' % (j, body[j][src].strip(), body[j][samp], body[j][ex]))
