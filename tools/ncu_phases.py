#!/usr/bin/env python
"""Per-phase view of an ncu source page: splits the SASS of one kernel at its BAR.SYNC instructions and
prints, per segment, warp instructions executed, stall samples and the dominant stall reasons.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_phases.py src.csv
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seg, segs = {"n": 0, "inst": 0, "samples": 0, "stalls": {}, "first": None, "shared_wf": 0, "shared_exc": 0}, []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    if seg["first"] is None:
        seg["first"] = r[ix["Address"]][-5:]
    seg["n"] += 1
    seg["inst"] += int(r[ix["Instructions Executed"]] or 0)
    seg["samples"] += int(r[ix["# Samples"]] or 0)
    seg["shared_wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    seg["shared_exc"] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v:
            seg["stalls"][c] = seg["stalls"].get(c, 0) + v
    if src.startswith("BAR.SYNC") or src.startswith("EXIT"):
        segs.append(seg)
        seg = {"n": 0, "inst": 0, "samples": 0, "stalls": {}, "first": None, "shared_wf": 0, "shared_exc": 0}
segs.append(seg)
tot_i = sum(s["inst"] for s in segs) or 1
tot_s = sum(s["samples"] for s in segs) or 1
print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
for k, s in enumerate(segs):
    top = sorted(s["stalls"].items(), key=lambda kv: -kv[1])[:4]
    print(f"seg {k:2d} @{s['first']} sass {s['n']:5d}  inst {s['inst']:>12,} ({100*s['inst']/tot_i:5.1f}%)  samples {s['samples']:>8,} ({100*s['samples']/tot_s:5.1f}%)"
          f"  smem wf {s['shared_wf']:>10,} (+{s['shared_exc']:,} excess)  " + " ".join(f"{c[6:]}:{v}" for c, v in top))
