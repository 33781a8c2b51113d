#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump over named source regions.
usage: ncu_regions.py dump.csv file:lo-hi=name [...]   (lines not covered by a region are listed under their file)"""
import csv
import sys
import collections

regions = []
for spec in sys.argv[2:]:
    loc, name = spec.split("=")
    f, rng = loc.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
rows = list(csv.reader(open(sys.argv[1])))
cur, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit() or len(r) <= 8 or r[2] != "-":
        continue
    ln = int(r[0])
    name = cur + ":other"
    for f, lo, hi, n in regions:
        if f == cur and lo <= ln <= hi:
            name = n
            break

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    a = agg.setdefault(name, [0.0, 0.0, 0.0])
    a[0] += num(r[6]); a[1] += num(r[7]); a[2] += num(r[8])
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  %s" % (100 * a[1] / ti, 100 * a[0] / ts, a[2] / a[1] if a[1] else 0, n))
