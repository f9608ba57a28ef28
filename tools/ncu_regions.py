#!/usr/bin/env python
"""Region breakdown of an `ncu --page source --csv` export: consecutive SASS instructions with the same execution
count are one region (a loop body, a straight-line block).  usage: ncu_regions.py prof_src.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
col = {k: i for i, k in enumerate(h)}
data = rows[2:]
ie, src, smp = col["Instructions Executed"], col["Source"], col["# Samples"]
tot = sum(int(r[ie]) for r in data)
ts = sum(int(r[smp]) for r in data)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
prev, start, acc, sacc, segs = None, 0, 0, 0, []
for k, r in enumerate(data):
    c = int(r[ie])
    if prev is None or abs(c - prev) > 0.02 * max(prev, 1):
        if prev is not None:
            segs.append((start, k - 1, prev, acc, sacc))
        start, acc, sacc = k, 0, 0
    prev = c
    acc += c
    sacc += int(r[smp])
segs.append((start, len(data) - 1, prev, acc, sacc))
print(f"total warp-instructions {tot:,}  samples {ts:,}")
for s in segs:
    if s[3] / tot > thr:
        print(f"instr {s[0]:5d}-{s[1]:5d} n={s[1]-s[0]+1:5d} exec/instr={s[2]:>12,} share={s[3]/tot:6.2%} samples={s[4]/max(ts,1):6.2%}  {data[s[0]][src][:48]}")
