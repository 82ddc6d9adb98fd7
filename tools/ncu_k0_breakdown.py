#!/usr/bin/env python
"""Source-level breakdown of an ncu report of k0_parse_kernel: stall samples / instructions per parser
region, the wavefront wait loop reported separately.  python tools/ncu_k0_breakdown.py rep.ncu-rep [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; h = None
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]
    elif r[0] == "Line No": h = r; ci = r.index("Instructions Executed"); cs = r.index("# Samples")
    elif h and r[0].isdigit() and len(r) > max(ci, cs):
        try:
            a = agg[(cur, int(r[0]))]; a[0] += int(r[cs]); a[1] += int(r[ci]); a[2] = r[1]
        except ValueError: pass
def is_wait(k, v):
    s = v[2]
    return "progress_load" in s or "backoff" in s or "nanosleep" in s or "ld.acquire" in s
wait = {k for k, v in agg.items() if is_wait(k, v)}
tw = sum(v[0] for k, v in agg.items()); ws = sum(agg[k][0] for k in wait); wi = sum(agg[k][1] for k in wait)
ts = tw - ws; ti = sum(v[1] for v in agg.values()) - wi
print("samples %d (wait loop %.1f%%), instructions outside the wait loop %d" % (tw, 100.0 * ws / max(tw, 1), ti))
for k, v in sorted(((k, v) for k, v in agg.items() if k not in wait), key=lambda x: -x[1][0])[:top]:
    print("%5.2f%% smp %5.2f%% ins | %s:%d | %s" % (100 * v[0] / ts, 100 * v[1] / ti, k[0][:12], k[1], v[2].strip()[:100]))
