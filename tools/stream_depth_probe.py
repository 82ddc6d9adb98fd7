import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/heif-decoder-lib_b200")
import bench, heif_b200 as hb
files = bench.make_content(8, os.path.join("/root/repo", "gpurun_out", "bench_content"))
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
