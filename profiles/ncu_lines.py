#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu_lines.py dump.csv [top]"""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("Function Name",) or hdr is None:
        continue
    if r[0].isdigit() and len(r) > 8 and r[2] == "-":  # per-source-line summary row
        key = (cur_file, int(r[0]))
        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        samples, inst, tinst = num(r[6]), num(r[7]), num(r[8])
        a = agg.setdefault(key, [0.0, 0.0, 0.0, r[1].strip()[:110]])
        a[0] += samples; a[1] += inst; a[2] += tinst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  thr/inst %4.1f  %s:%d  %s" % (100 * a[0] / tot_s, 100 * a[1] / tot_i, a[2] / a[1] if a[1] else 0, f, ln, a[3]))
