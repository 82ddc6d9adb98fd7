import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/heif-decoder-lib_b200"); sys.path.insert(0, "/root/repo/tests")
import bench
files = bench.make_content(2, os.path.join("/root/repo", "gpurun_out", "bench_content"))
for k in range(2):
    r = bench.plugin_arm(files, os.cpu_count())
    print("plugin", round(r["value"], 1), "libde265", round(r["libde265_plugin_same_call"]["value"], 1), r["bit_exact_vs_libde265_plugin"])
