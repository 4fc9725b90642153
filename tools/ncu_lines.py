#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass,cuda` dump per CUDA source line."""
import csv, sys, collections
path = sys.argv[1]; per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ix = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
agg = []
for r in rows:
    if len(r) <= ix or r[0] in ("", "Line No"): continue
    try: agg.append((int(r[0]), r[1], int(r[ix]), int(r[isamp])))
    except ValueError: pass
tot = sum(a[2] for a in agg); ts = sum(a[3] for a in agg)
print(f"total instr {tot/per:.1f}  samples {ts}")
for ln, src, n, s in sorted(agg, key=lambda a: -a[2])[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print(f"{ln:5d} {n/per:9.1f} {100*n/tot:5.1f}%  samp {100*s/ts:5.1f}%  {src.strip()[:90]}")
