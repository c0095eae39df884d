#!/usr/bin/env python
"""Per-source-line warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`:
where the resident warps SPEND their time (samples), next to where instructions are issued.
usage: ncu_stalls.py file.csv [min_pct]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None; samp = collections.OrderedDict(); src = {}; inst = {}; why = {}
for r in rows:
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; cols = {k: i for i, k in enumerate(r)}; continue
    if hdr is None or len(r) < 48 or not r[0].strip().isdigit() or r[2] != "-":
        continue
    ln = int(r[0]); src[ln] = r[1]
    def f(x):
        try: return int(float(x))
        except ValueError: return 0
    samp[ln] = samp.get(ln, 0) + f(r[6]); inst[ln] = inst.get(ln, 0) + f(r[7])
    w = why.setdefault(ln, collections.Counter())
    for i in range(31, 48):
        w[hdr[i]] += f(r[i])
tot = sum(samp.values()) or 1
print("total samples", tot)
for ln, v in samp.items():
    if 100.0 * v / tot >= minp:
        top = ", ".join("%s %d" % (k.replace("stall_", ""), c) for k, c in why[ln].most_common(3) if c)
        print("%5d %7d %5.1f%%  inst %9d  [%s]  %s" % (ln, v, 100.0 * v / tot, inst[ln], top, src[ln].strip()[:80]))
