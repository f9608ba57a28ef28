#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: executed-instruction histogram by opcode, stall samples, and the
hottest contiguous regions.  usage: ncu_src_summary.py prof_src.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
col = {k: i for i, k in enumerate(h)}
data = rows[2:]
ie, src, smp = col["Instructions Executed"], col["Source"], col["# Samples"]
te = col["Thread Instructions Executed"]
tot = sum(int(r[ie]) for r in data)
tsamp = sum(int(r[smp]) for r in data)
ops = collections.Counter()
samp = collections.Counter()
thr = collections.Counter()
for r in data:
    op = r[src].split()[0]
    if op.startswith("@"):
        op = r[src].split()[1]
    op = op.split(".")[0]
    ops[op] += int(r[ie]); samp[op] += int(r[smp]); thr[op] += int(r[te])
print(f"total warp-instructions {tot:,}  samples {tsamp:,}  static instrs {len(data)}")
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"  {op:12s} {c:>16,} {c/tot:7.3%}  samples {samp[op]/max(tsamp,1):7.3%}  lanes {thr[op]/max(c,1):5.1f}")
