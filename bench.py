#!/usr/bin/env python
"""bench.py — decoded MP/s HEIC -> RGB on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (config.workload): BASELINE config C2 — 4032x3024 "iPhone-style" grid of 48 (8x6) 512x512
HEVC intra tiles, 8-bit 4:2:0, CTB 64, WPP, SAO + deblocking, QP 26, full-range BT.601 VUI -> interleaved
RGB. Content is synthetic (tools/hevc_enc closed-loop encoder + tools/heif_writer), generated untimed
at start-up. One step = one batch of `--images` such files (default 32 = 1536 coded pictures in flight).

  value : device time of K steps of K1..K5 with the packed records already resident in HBM
          (CUDA events on the engine's stream), whole-job MP/s over all ranks
  e2e   : the same metric through the C ABI (hc_heic_decode_stream) from HEIC bytes in host memory to RGB bytes in
          pinned host memory: container + header parse, slice-data parse (K0 on the GPU, plus the share the host
          threads take meanwhile; --parser host: host threads only), H2D, K1..K5 and D2H inside the timed region
  roofline     : the dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : the unmodified reference (oracle/_ref: libheif + libde265, heif_decode_image -> RGB) on
                 the host cores, bounded sample of the same file

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
Launched under torchrun for N > 1 (one rank per GPU, NCCL only for the barrier / max-over-ranks).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "heif-decoder-lib_b200"), os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOAD = "C2: 4032x3024 grid of 48 512x512 HEVC-intra tiles, 8-bit 4:2:0, CTB64 WPP SAO+deblock QP26 -> RGB24"
GRID_W, GRID_H, TILE = 4032, 3024, 512
METRIC = "decoded MP/s HEIC->RGB (device-timed)"
UNIT = "MP/s"


def make_content(n_distinct, cache_dir):
    """n_distinct synthetic 12 MP grid HEICs (cached on disk between runs of the same box)."""
    from tools import heif_writer
    os.makedirs(cache_dir, exist_ok=True)
    files = []
    for i in range(n_distinct):
        path = os.path.join(cache_dir, "c2_%dx%d_t%d_seed%d.heic" % (GRID_W, GRID_H, TILE, i))
        if not os.path.exists(path):
            data = heif_writer.synth_grid_heic(GRID_W, GRID_H, tile=TILE, seed=100 + i, qp=26, wpp=1, sao=1, log2_ctb=6)
            with open(path + ".tmp", "wb") as f:
                f.write(data)
            os.replace(path + ".tmp", path)
        files.append(open(path, "rb").read())
    return files


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def reference_arm(files, steps, warmup, threads):
    """The reference's own CPU implementation on all host cores, two ways, the better one is the value:
      (a) its own policy: heif_decode_image(..., RGB, interleaved_RGB) per file with
          heif_context_set_threads(ctx, handle, cores) (grid -> one thread per tile, heif.cc:499-514);
      (b) process-level data parallelism: `cores` files decoded concurrently, one thread each."""
    import concurrent.futures as cf
    import refheif as R
    if not R.available():
        return None
    mp = GRID_W * GRID_H / 1e6
    f0 = files[0]
    for _ in range(warmup):
        R.decode(f0, R.COLORSPACE_RGB, R.CHROMA_RGB, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        R.decode(f0, R.COLORSPACE_RGB, R.CHROMA_RGB, threads=threads)
    dt_a = (time.perf_counter() - t0) / steps
    with cf.ThreadPoolExecutor(max_workers=threads) as pool:   # ctypes releases the GIL inside heif_decode_image
        def one(_):
            R.decode(f0, R.COLORSPACE_RGB, R.CHROMA_RGB, threads=1)
        list(pool.map(one, range(threads)))                    # warm-up
        t0 = time.perf_counter()
        n = threads * max(1, min(steps, 2))
        list(pool.map(one, range(n)))
        dt_b = (time.perf_counter() - t0) / n
    best = min(dt_a, dt_b)
    return {"value": mp / best, "ms_per_step": best * 1e3,
            "sample": "one 12.19 MP grid file: (a) %d tile threads %.1f MP/s, (b) %d concurrent single-thread decodes %.1f MP/s" % (
                threads, mp / dt_a, threads, mp / dt_b)}


def plugin_arm(files, threads):
    """Level-1 drop-in: the UNMODIFIED reference libheif (oracle/_ref) decodes the same file through its own
    heif_decode_image, with libheif-cuda.so loaded by heif_load_plugin and selected as decoder ("cuda"): host parse +
    K1..K4 per tile inside the plugin, grid paste + colour conversion by libheif on the CPU. Checked against the
    reference's libde265 plugin on the same file."""
    import ctypes as C
    import hashlib
    import refheif as R
    plugin = os.path.join(ROOT, "heif-decoder-lib_b200", "plugins", "libheif-cuda.so")
    if not R.available() or not os.path.exists(plugin):
        return None
    try:
        L = R.lib()
        os.environ["HEIFCUDA_LIBHEIF"] = os.path.join(R.REF_DIR, "libheifref.so")
        info = C.c_void_p()
        err = L.heif_load_plugin(plugin.encode(), C.byref(info))
        if err.code != 0:
            return {"error": (err.message or b"").decode()}
        R.FOREIGN_PLUGIN_LOADED = True
        f0 = files[0]
        mp = GRID_W * GRID_H / 1e6
        want = R.decode(f0, R.COLORSPACE_RGB, R.CHROMA_RGB, threads=threads, decoder_id="libde265")["interleaved"][0]
        got = R.decode(f0, R.COLORSPACE_RGB, R.CHROMA_RGB, threads=threads, decoder_id="cuda")["interleaved"][0]   # also the warm-up
        n = 3
        t0 = time.perf_counter()
        for _ in range(n):
            R.decode(f0, R.COLORSPACE_RGB, R.CHROMA_RGB, threads=threads, decoder_id="cuda")
        dt = (time.perf_counter() - t0) / n
        return {"value": mp / dt, "unit": UNIT, "ms_per_image": dt * 1e3, "tile_threads": threads,
                "bit_exact_vs_libde265_plugin": hashlib.md5(got).digest() == hashlib.md5(want).digest(),
                "what": "unmodified reference libheif, heif_decode_image -> RGB, decoder plugin = libheif-cuda.so (one file at a time; "
                        "grid paste and colour conversion stay on the CPU inside libheif)"}
    except Exception as e:   # never fail the measurement on the optional arm
        return {"error": str(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--images", type=int, default=32, help="12 MP files per step and GPU (K0 is a wavefront per picture: its ramps amortise better over 32 files than 16)")
    ap.add_argument("--distinct", type=int, default=2, help="distinct synthetic files (replicated to --images)")
    ap.add_argument("--threads", type=int, default=0, help="host parse threads (0 = all cores)")
    ap.add_argument("--host-share", type=int, default=-1, help="with --parser device: %% of the coded items parsed by the host threads "
                    "meanwhile in the e2e arm (default: engine default)")
    ap.add_argument("--parser", default="device", choices=["device", "host"],
                    help="where the CABAC slice data is parsed: K0 on the GPU (default) or the host parser")
    ap.add_argument("--skip-baselines", action="store_true", help="kernel experiments: leave out the cpu_baseline and plugin_dropin arms")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    cache = os.path.join(ROOT, "gpurun_out", "bench_content")

    if args.impl == "reference":
        if rank != 0:
            return 0
        files = make_content(1, cache)
        r = reference_arm(files, max(1, min(args.steps, 5)), 1, cores)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference build) is missing"}))
            return 0
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": {"workload": WORKLOAD},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "reference", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # stdout carries exactly one JSON line: NCCL's version / debug lines go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import numpy as np
    import torch
    import heif_b200 as hb

    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the reconstruction path has no CPU fallback")
    torch.cuda.set_device(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- content (untimed; rank 0 generates, the others read the cache) ----
    if rank == 0:
        distinct = make_content(args.distinct, cache)
    barrier()
    if rank != 0:
        distinct = make_content(args.distinct, cache)
    files = [distinct[(i + rank) % len(distinct)] for i in range(args.images)]
    mp_per_step = args.images * GRID_W * GRID_H / 1e6

    eng = hb.Engine(local_rank)
    eng.set_option("device_parse", 1 if args.parser == "device" else 0)
    if args.host_share >= 0:
        eng.set_option("host_share_pct", args.host_share)
    threads = args.threads or cores // max(1, world)
    # R: bytes of packed records per output pixel as the host parser emits them (what K1/K2 read from HBM either way)
    hf = hb.HeifFile(files[0], host_only=False)
    rec_bytes_per_px = hb.parse_picture(hf.coded_stream(hf.grid_tiles(hf.primary_id)[0])).upload_bytes / float(TILE * TILE)
    hf.close()

    # ---- device-resident arm: records in HBM, K1..K5 timed with CUDA events on the engine stream ----
    job = hb.HeicJob(eng, files, want_alpha=False, threads=threads)
    job.upload()
    for _ in range(args.warmup):
        job.run()
    job.sync()
    stage = job.stage_ms()   # per-kernel split of the last warm-up step
    # clocks are sampled over >= 0.6 s of the same load (extra untimed steps) followed by the timed steps:
    # K steps of a few ms each are too short for nvidia-smi's polling on their own
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms = max(1e-3, sum(stage[k] for k in ("k1_transform", "k2_intra", "k3_deblock", "k4_sao", "k5_csc")))
    for _ in range(int(600.0 / step_ms) + 1):
        job.run()
    job.sync()
    barrier()
    job.timer_start()
    for _ in range(args.steps):
        job.run()
    dev_ms = job.timer_stop_ms()
    barrier()
    clocks = sampler.stop()
    dev_ms = max_over_ranks(dev_ms)
    launches = job.launch_count * args.steps
    upload_bytes = job.upload_bytes
    rgb_bytes = sum(d.width * d.height * d.bytes_per_pixel for d in job.descs)
    stage_last = job.stage_ms()

    # ---- parity guard: the timed configuration still produces the reference's pixels ----
    check = None
    try:
        import hashlib
        import heic_oracle
        got = job.read_rgb(0)
        want = heic_oracle.decode_rgb(files[0], hb.OUT_RGB)
        check = bool(np.array_equal(got, want))
    except Exception as e:  # the oracle is only the checker; never fail the measurement on it
        check = "unchecked: %s" % e
    job.close()

    # ---- end-to-end arm: HEIC bytes (host) -> RGB bytes (pinned host), everything inside the timed region ----
    # One call of the public streaming API (hc_heic_decode_stream) over (1 warm-up + e2e_steps) batches of
    # `--images` files: the host CABAC parse of batch b+1 overlaps H2D + kernels + D2H of batch b, which is how a
    # long file list is meant to be fed (BASELINE config C4). The first batch (pipeline fill, allocations) is
    # timed separately and excluded.
    e2e_steps = max(6, min(args.steps, 16))
    marks = []
    checksum = [0]

    def on_image(index, desc, rows):
        if index % args.images == args.images - 1:
            marks.append(time.perf_counter())
        if index == 0:
            checksum[0] = int(rows[::97, ::389].astype(np.uint64).sum())   # touch the pinned result on the host

    barrier()
    t_start = time.perf_counter()
    st = hb.decode_stream(eng, files * (1 + e2e_steps), on_image, want_alpha=False, threads=threads, files_per_batch=args.images)
    barrier()
    e2e_dt = max_over_ranks((marks[-1] - marks[0]) / e2e_steps)
    parse_s = st["seconds_parse"] / st["batches"]
    gpu_phase_s = st["seconds_gpu_phase"] / st["batches"]
    first_batch_s = marks[0] - t_start

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    px = args.images * GRID_W * GRID_H
    coded_px = args.images * 48 * TILE * TILE
    # algorithmic bytes per kernel for 8-bit 4:2:0 (DESIGN.md "kernels")
    alg = {"k1_transform": rec_bytes_per_px * coded_px + 3.0 * coded_px, "k2_intra": 4.5 * coded_px, "k3_deblock": 2 * 3.0 * coded_px,
           "k4_sao": 1.5 * coded_px + 1.5 * px, "k5_csc": 4.5 * px}
    kernels = {k: stage_last[k] for k in alg}
    k0_ms = stage_last.get("k0_parse", 0.0)
    dom = max(kernels, key=kernels.get)
    achieved = alg[dom] / (kernels[dom] * 1e-3) / 1e9 if kernels[dom] > 0 else 0.0
    # DRAM traffic of the dominant kernel per launch, from the committed ncu capture of this very configuration
    traffic, traffic_src = None, None
    try:
        entries = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(dom) or []
        for t in (entries if isinstance(entries, list) else [entries]):
            if t["images_per_step"] == args.images:
                traffic, traffic_src = t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write, ncu)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg[dom], "peak_source": peak_src,
                "all_kernels": {k: {"ms": round(kernels[k], 4), "algorithmic_GBps": round(alg[k] / (kernels[k] * 1e-3) / 1e9, 1) if kernels[k] > 0 else None}
                                for k in alg},
                "k0_parse": {"ms": round(k0_ms, 4), "note": "device CABAC parse, once per upload, outside `value`, inside e2e; serial per substream: "
                             "bound by single-thread instruction latency, not by HBM (DESIGN.md)"} if k0_ms > 0 else None,
                "note": "k2_intra is a dependency-latency-bound CTB wavefront, not a streaming kernel (DESIGN.md)"}

    # ---- CPU baseline on the box's cores (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.skip_baselines:
        r = reference_arm(distinct, 2, 1, cores)
        if r is not None:
            cpu = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "reference", "sample": r["sample"]}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "oracle/_ref missing on this box"}

    plugin = None
    if rank == 0 and world == 1 and not args.skip_baselines:
        eng.close()
        eng = None
        plugin = plugin_arm(distinct, cores)

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        line = {
            "metric": METRIC, "value": world * mp_per_step / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_step_per_gpu": args.images, "coded_pictures_per_step_per_gpu": args.images * 48,
                       "l2": "working set per step (planes + residuals + RGB, ~%d MB) exceeds the 126 MB L2; no explicit flush" % (
                           args.images * (18 + 18 + 37 + 38)),
                       "host_parse_threads": threads, "parity_vs_oracle": check},
            "e2e": {"value": world * mp_per_step / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": int(st["bytes_h2d"] / max(1, st["batches"])),
                    "d2h_bytes_per_step": int(st["bytes_d2h"] / max(1, st["batches"])),
                    "ms_per_step": e2e_dt * 1e3, "host_parse_ms_per_step": parse_s * 1e3, "gpu_phase_ms_per_step": gpu_phase_s * 1e3,
                    "first_batch_ms": first_batch_s * 1e3, "steps": e2e_steps,
                    "api": "hc_heic_decode_stream: three batches in flight on the GPU (K0 on a low-priority stream, K1..K5 + copies on "
                           "high-priority ones), header parse (+ host share of the slice data) two batches ahead; pinned host output"},
            "gpu_launches": launches,
            "clocks": clocks,
            "stage_ms_last_step": {k: round(v, 4) for k, v in stage_last.items()},
            "value_incl_parse": (world * mp_per_step / ((ms_per_step + stage_last.get("k0_parse", 0.0)) * 1e-3)) if args.parser == "device" else None,
            "record_bytes_per_px": rec_bytes_per_px, "uploaded_bytes_per_px": upload_bytes / px,
            "parser": "K0 on the device (runs once per upload; `value` times K1..K5 with the records resident in HBM, k0_parse is "
                      "listed in stage_ms_last_step and is inside e2e)" if args.parser == "device" else "host CABAC parser",
            "roofline": roofline,
            "cpu_baseline": cpu,
            "plugin_dropin": plugin,
        }
        print(json.dumps(line))
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
