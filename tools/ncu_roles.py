#!/usr/bin/env python
"""Stall samples and executed instructions of a kernel split at its named barriers / bulk copies / exits (SASS order), i.e. per
warp role of a warp-specialised kernel.   ncu -i x.ncu-rep --page source --csv --print-source sass > x.csv; python tools/ncu_roles.py x.csv"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[1]
isrc=h.index("Source"); isamp=h.index("# Samples"); iex=h.index("Instructions Executed")
st=[c for c in h if c.startswith("stall_") and "Not Issued" not in c]
idx={c:h.index(c) for c in st}
seg_s=0; seg_e=0; segst=collections.Counter(); total=0; n=0
segs=[]
for r in rows[2:]:
    if len(r)<len(h): continue
    s=int(r[isamp]); e=int(r[iex]); total+=s
    seg_s+=s; seg_e+=e; n+=1
    for c in st: segst[c]+=int(r[idx[c]] or 0)
    src=r[isrc].strip()
    op=src.split()[1] if src.startswith("@") else src.split()[0]
    if op.startswith("BAR") or op.startswith("EXIT") or "UBLKCP" in op:
        if seg_s>0: segs.append((n,seg_s,seg_e,src[:50],segst.most_common(4)))
        seg_s=0; seg_e=0; segst=collections.Counter()
print("total samples",total)
for i,(n,s,e,src,top) in enumerate(segs):
    print("%5d %6d %5.1f%% ex=%8d  %-50s %s"%(n,s,100*s/total,e,src,[(c[6:],v) for c,v in top]))
