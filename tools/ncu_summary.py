#!/usr/bin/env python
"""Summarises an ncu report (read on the CPU box): headline counters per kernel launch and, with
--source, the hottest CUDA source lines by stall samples (needs -lineinfo + --import-source on).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source] [--top 40]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  (id %s)" % (d.get("Kernel Name", "?")[:90], d.get("ID")))
        for k in KEYS:
            if k in d:
                print("   %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
        stalls = [(k, float(v)) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v]
        for k, v in sorted(stalls, key=lambda x: -x[1])[:8]:
            print("   stall %-64s %.2f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))


def source(rep, top):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    cur, h, ci, cs = None, None, 0, 0
    agg = collections.defaultdict(lambda: [0, 0, ""])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            h, ci, cs = r, r.index("Instructions Executed"), r.index("# Samples")
        elif h and r[0].isdigit() and len(r) > max(ci, cs):
            try:
                agg[(cur, int(r[0]))][0] += int(r[cs])
                agg[(cur, int(r[0]))][1] += int(r[ci])
                agg[(cur, int(r[0]))][2] = r[1]
            except ValueError:
                pass
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print("stall samples %d, warp instructions %d" % (ts, ti))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print("%6.2f%% smp %6.2f%% ins | %s:%d | %s" % (100.0 * v[0] / ts, 100.0 * v[1] / ti, k[0][:14], k[1], v[2].strip()[:100]))


if __name__ == "__main__":
    rep = sys.argv[1]
    raw(rep)
    if "--source" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
        source(rep, top)
