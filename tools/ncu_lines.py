#!/usr/bin/env python
"""Per-source-line instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu_lines.py file.csv [min_pct]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdr = None; agg = collections.OrderedDict(); src = {}; wf = {}
for r in rows:
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 21 or not r[0].strip().isdigit() or r[2] != "-":
        continue
    ln = int(r[0]); src[ln] = r[1]
    try:
        agg[ln] = agg.get(ln, 0) + int(float(r[7])); wf[ln] = wf.get(ln, 0) + int(float(r[19] or 0))
    except ValueError:
        pass
tot = sum(agg.values()) or 1
print("total warp instructions", tot, " shared wavefronts", sum(wf.values()))
for ln, v in agg.items():
    if 100.0 * v / tot >= minp:
        print("%5d %12d %5.1f%%  wf %10d  %s" % (ln, v, 100.0 * v / tot, wf[ln], src[ln].strip()[:100]))
