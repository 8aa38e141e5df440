#!/usr/bin/env python3
"""Stall samples and executed instructions per CUDA source line of an ncu report.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python profiles/src_hot.py [min_pct] [file-substring]"""
import csv, sys
minpct = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
want = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(sys.stdin))
def f(x):
    try: return int(float(x))
    except ValueError: return 0
cur_file = ""
per = {}   # (file, line) -> [samples, exec, source]
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        cur_file = r[1]; continue
    if len(r) < 8 or r[0] == "Line No": continue
    line, src, addr = r[0], r[1], r[2]
    if not line.isdigit(): continue
    key = (cur_file, int(line))
    e = per.setdefault(key, [0, 0, src])
    if addr:   # a SASS row attributed to this source line
        e[0] += f(r[4]); e[1] += f(r[7])
    else:
        e[2] = src
tot = sum(v[0] for v in per.values()) or 1
print("total samples", tot)
for (fn, ln), v in sorted(per.items()):
    if want and want not in fn: continue
    if 100.0 * v[0] / tot >= minpct:
        print("%-14s %4d %5.1f%% exec=%10d | %s" % (fn.split('/')[-1], ln, 100.0 * v[0] / tot, v[1], v[2].strip()[:100]))
