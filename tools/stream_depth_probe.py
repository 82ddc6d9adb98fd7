"""Steady state of hc_heic_decode_stream over N batches of 64 twelve-megapixel files: whole call, first-to-last delivery and
the middle half of the run (the figure that does not depend on pipeline fill and drain). HEIFCUDA_STREAM_DEPTH=3..6 fixes
the batches in flight.   python tools/stream_depth_probe.py 40"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heif-decoder-lib_b200")):
    sys.path.insert(0, p)
import bench, heif_b200 as hb
files = bench.make_content(8, os.path.join(ROOT, "gpurun_out", "bench_content"))
eng = hb.Engine(0)
n, per = int(sys.argv[1]), 64
lst = [files[i % 8] for i in range(per)] * n
marks = []
def on_image(i, d, rows):
    if i % per == per - 1: marks.append(time.perf_counter())
hb.decode_stream(eng, lst[:per * 3], None, threads=16, files_per_batch=per)   # warm allocations
t0 = time.perf_counter()
st = hb.decode_stream(eng, lst, on_image, threads=16, files_per_batch=per)
t1 = time.perf_counter()
mp = per * 12.192768
print("depth", os.environ.get("HEIFCUDA_STREAM_DEPTH", "3"), "batches", n, "whole call: %.1f ms/batch = %.0f MP/s" % ((t1 - t0) / n * 1e3, mp * n / (t1 - t0)),
      "| first->last delivery: %.1f ms/batch = %.0f MP/s" % ((marks[-1] - marks[0]) / (n - 1) * 1e3, mp * (n - 1) / (marks[-1] - marks[0])),
      "| middle half: %.1f ms/batch" % ((marks[3 * n // 4] - marks[n // 4]) / (3 * n // 4 - n // 4) * 1e3))
