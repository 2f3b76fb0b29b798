#!/usr/bin/env python3
"""stdin: `ncu --page source --csv` of one kernel; stdout: the source lines with the most warp-stall samples (needs -lineinfo)"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = {h: j for j, h in enumerate(r)}; start = i + 1; break
if hdr is None:
    print("no source page"); sys.exit(0)
samp = next((h for h in hdr if h.startswith("Warp Stall Sampling (All")), None) or next(h for h in hdr if "Sampl" in h)
inst = next((h for h in hdr if h.startswith("Instructions Executed")), None)
items = []
for r in rows[start:]:
    try:
        items.append((float(r[hdr[samp]] or 0), r[hdr["Source"]][:150], r[hdr[inst]] if inst else ""))
    except (ValueError, IndexError):
        continue
tot = sum(i[0] for i in items) or 1.0
print(f"# top source lines by {samp} (total {tot:.0f})")
for s, src, n in sorted(items, key=lambda t: -t[0])[:45]:
    print(f"{100 * s / tot:5.1f}%  inst={n:>10s}  {src}")
