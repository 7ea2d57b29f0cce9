#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`).
    python tools/launch_summary.py launches.csv [-o summary.md] [--title "..."]"""
import collections
import csv
import sys

src = sys.argv[1]
dst = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else f"ncu launch list `{src}`"
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("nsecond", "ns") else v * 1e3 if unit in ("msecond", "ms") else v
    k = r[ix["Kernel Name"]]
    tot[k][0] += 1
    tot[k][1] += us
allus = sum(v[1] for v in tot.values())
lines = [f"# {title}", "", "| launches | avg (us) | share of all captured GPU time | kernel |", "|---:|---:|---:|---|"]
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:14]:
    lines.append(f"| {n} | {us / n:.1f} | {100 * us / allus:.2f} % | `{k[:110]}` |")
text = "\n".join(lines)
if dst:
    open(dst, "w").write(text + "\n")
print(text)
