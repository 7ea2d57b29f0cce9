#!/usr/bin/env python
"""Run-length view of one segment of an ncu source page: consecutive SASS instructions with the same execution count.
    python tools/ncu_runs.py src.csv <first-address-suffix> <last-address-suffix>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
a0, a1 = sys.argv[2], sys.argv[3]
start = False; prev = None; n = 0; smp = 0; first = ""
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = r[ix['Address']][-5:]
    if a == a0: start = True
    if a == a1: break
    if start:
        c = int(r[ix['Instructions Executed']])
        if c != prev:
            if prev is not None: print(f"{n:4d} instrs x {prev:>9,} = {n*prev:>12,}  samples {smp:6d}  first: {first}")
            prev = c; n = 0; smp = 0; first = a + ' ' + r[ix['Source']].strip()[:60]
        n += 1; smp += int(r[ix['# Samples']])
print(f"{n:4d} instrs x {prev:>9,} = {n*prev:>12,}  samples {smp:6d}  first: {first}")
