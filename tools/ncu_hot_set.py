#!/usr/bin/env python
"""Instruction-cache view of an ncu report (--import-source on): how many static SASS instructions account for 50 / 80 /
90 / 95 / 99 % of the executed instructions — the number that mattered for K0.
  python tools/ncu_hot_set.py rep.ncu-rep"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h=None; ins=[]
for r in rows:
    if not r: continue
    if r[0]=="Address" or (h is None and "Instructions Executed" in r):
        h=r; ci=r.index("Instructions Executed"); cs=r.index("Source") if "Source" in r else 1
    elif h and len(r)>ci:
        try: ins.append((int(r[ci]), r[0], r[cs]))
        except ValueError: pass
print(len(ins),"sass instructions")
tot=sum(i[0] for i in ins)
s=sorted(ins,reverse=True)
acc=0
for n,(c,a,src) in enumerate(s):
    acc+=c
    for q in (0.5,0.8,0.9,0.95,0.99):
        if acc-c < q*tot <= acc: print("%d%% of executed instructions come from %d static instructions (%.1f KB)"%(q*100,n+1,(n+1)*16/1024))
