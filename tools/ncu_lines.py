#!/usr/bin/env python
"""Stall samples and executed instructions per source line of an ncu report (--import-source on, -lineinfo build).
  python tools/ncu_lines.py rep.ncu-rep [top]"""
import collections, csv, io, subprocess, sys
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
cur=None;h=None
agg=collections.defaultdict(lambda:[0,0,""])
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]
    elif r[0]=="Line No": h=r; ci=r.index("Instructions Executed"); cs=r.index("# Samples")
    elif h and r[0].isdigit() and len(r)>max(ci,cs):
        try:
            a=agg[(cur,int(r[0]))]; a[0]+=int(r[cs]); a[1]+=int(r[ci]); a[2]=r[1]
        except ValueError: pass
ts=sum(v[0] for v in agg.values()); ti=sum(v[1] for v in agg.values())
print("samples",ts,"instructions",ti)
for k,v in sorted(agg.items(), key=lambda x:-x[1][0])[:top]:
    print("%5.2f%% smp %5.2f%% ins | %s:%d | %s"%(100*v[0]/ts,100*v[1]/ti,k[0][:14],k[1],v[2].strip()[:105]))
