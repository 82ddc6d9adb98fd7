"""The plugin arm of bench.py alone (fresh process, nothing else on the GPU): libheif-cuda.so vs the libde265 plugin through the
unmodified reference libheif, median of seven single-file decodes each."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heif-decoder-lib_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench
files = bench.make_content(2, os.path.join(ROOT, "gpurun_out", "bench_content"))
for k in range(2):
    r = bench.plugin_arm(files, os.cpu_count())
    print("plugin", round(r["value"], 1), "libde265", round(r["libde265_plugin_same_call"]["value"], 1), r["bit_exact_vs_libde265_plugin"])
