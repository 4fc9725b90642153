#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump: per CUDA source line
(samples, instructions, dominant stall reasons) and per SASS opcode."""
import csv, sys, collections
path = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0   # neighbourhoods per launch
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ia, isrc = hdr.index("Address"), hdr.index("Address") + 1
ins, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by_line = collections.defaultdict(lambda: collections.Counter())
by_op = collections.defaultdict(lambda: collections.Counter())
tot_s = tot_i = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) - 2:
        continue
    try:
        n, s = int(r[ins]), int(r[isamp])
    except ValueError:
        continue
    line = r[0]
    sass = r[isrc].strip()
    op = sass.split()[0] if sass else "?"
    if op.startswith("@"):
        op = sass.split()[1]
    op = op.split(".")[0] + ("." + sass.split()[0].split(".")[1] if op in ("MUFU",) and "." in sass.split()[0] else "")
    tot_s += s; tot_i += n
    for d in (by_line[(line, r[1].strip()[:80])], by_op[op]):
        d["instr"] += n; d["samples"] += s
        for i in stall:
            d[hdr[i]] += int(r[i] or 0)
print(f"instructions/nbhd {tot_i/per:.0f}  samples {tot_s}")
def show(d, top):
    for key, v in sorted(d.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((k, c) for k, c in v.items() if k.startswith("stall_")), key=lambda kc: -kc[1])[:3]
        sts = " ".join(f"{k[6:]}={100*c/max(v['samples'],1):.0f}%" for k, c in st)
        print(f"{str(key)[:100]:100s} instr/nbhd {v['instr']/per:7.1f} samp {100*v['samples']/tot_s:5.1f}%  {sts}")
print("== by opcode"); show(by_op, 30)
print("== by source line"); show(by_line, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
