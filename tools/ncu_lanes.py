#!/usr/bin/env python
"""Per source line of an ncu report: share of the executed warp instructions, average active lanes per instruction and
stall samples — the view that matters for the lockstep K0 (how many lanes share a handler).
  python tools/ncu_lanes.py rep.ncu-rep [top] [file-substring]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40; only = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
cur = None; h = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]
    elif r[0] == "Line No": h = r; ci = r.index("Instructions Executed"); ct = r.index("Thread Instructions Executed"); cs = r.index("# Samples")
    elif h and r[0].isdigit() and len(r) > max(ci, ct, cs):
        try:
            a = agg[(cur, int(r[0]))]; a[0] += int(r[ci]); a[1] += int(r[ct]); a[2] += int(r[cs]); a[3] = r[1]
        except ValueError: pass
tot = sum(v[0] for v in agg.values()); tt = sum(v[1] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print("warp instructions %d, avg active lanes %.2f, samples %d" % (tot, tt / max(tot, 1), ts))
for k, v in sorted(((k, v) for k, v in agg.items() if only in (k[0] or "")), key=lambda x: -x[1][0])[:top]:
    print("%5.2f%% ins %5.2f lanes %5.2f%% smp | %s:%d | %s" % (100 * v[0] / tot, v[1] / max(v[0], 1), 100 * v[2] / max(ts, 1), (k[0] or "?")[:14], k[1], v[3].strip()[:80]))
