#!/usr/bin/env python
"""Aggregate an ncu SASS-level source CSV (--page source --csv) by address range -> stall reasons."""
import csv, sys, re, collections
path=sys.argv[1]
ranges=[(n,int(a,16),int(b,16)) for n,a,b in (x.split(':') for x in sys.argv[2:])]
rows=list(csv.reader(open(path)))
hdr=rows[1]
ia=hdr.index("Address"); ins=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples")
stall=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base=None
res={n:collections.Counter() for n,_,_ in ranges}
for r in rows[2:]:
    if len(r)<len(hdr)-2: continue
    try: addr=int(r[ia],16)
    except: continue
    if base is None: base=addr
    off=addr-base
    for n,a,b in ranges:
        if a<=off<b:
            res[n]['instr']+=int(r[ins]); res[n]['samples']+=int(r[isamp])
            for i in stall: res[n][hdr[i]]+=int(r[i] or 0)
tot=sum(v['samples'] for v in res.values())
for n,v in res.items():
    print(f"== {n}: instr {v['instr']} samples {v['samples']} ({100*v['samples']/max(tot,1):.1f}%)")
    for k,c in v.most_common(12):
        if k.startswith('stall_'): print(f"     {k:24s} {100*c/max(v['samples'],1):5.1f}%")
