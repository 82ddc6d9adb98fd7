"""Damaged slice data must never crash, hang or corrupt memory — only decode to (wrong) pixels or be reported.

CPU: the host CABAC parser and the K0 core (run on the CPU) on seeded corruptions of several fixtures, in a child
process so that a crash would fail the test instead of the test run (libde265 has the same exposure: the plugin feeds
whatever libheif hands it). GPU: the same corruptions as ONE K0 batch together with undamaged pictures, which must
still decode bit-exactly (a failing chain releases its wavefront waiters and reports; nothing else is disturbed)."""
import glob
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import heif_b200 as hb
from conftest import ROOT

GEN_DIR = os.path.join(ROOT, "tests", "golden", "generated")
GEN_META = json.load(open(os.path.join(ROOT, "tests", "golden", "generated.json")))
VICTIMS = ["base_420_8", "wpp", "wpp_slices", "slices_dep", "c422_10", "cuqpdelta_d2", "scaling_custom", "tskip", "tile_512", "qp10"]


def _victims():
    have = {os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GEN_DIR, "*.hevc"))}
    v = [n for n in VICTIMS if n in have]
    return v if v else sorted(have)[:6]


def _fmt(data):
    return hb.STREAM_ANNEXB if data[:3] == b"\x00\x00\x01" or data[:4] == b"\x00\x00\x00\x01" else hb.STREAM_LENGTH_PREFIXED


def corrupt(data, seed):
    """Flips / overwrites bytes in the last two thirds of the stream (slice data), avoiding new start codes."""
    rng = np.random.default_rng(seed)
    d = bytearray(data)
    lo = len(d) // 3
    for _ in range(int(rng.integers(1, 12))):
        i = int(rng.integers(lo, len(d)))
        n = int(rng.integers(1, 24))
        for k in range(i, min(len(d), i + n)):
            v = int(rng.integers(1, 256))
            d[k] = v if v > 3 else 0x80
    if rng.integers(0, 4) == 0:
        d = d[:int(rng.integers(lo, len(d)))]          # truncation
    return bytes(d)


_CHILD = r"""
import sys, os
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import heif_b200 as hb
import test_fuzz as T
ok = rejected = 0
for name in T._victims():
    data = open(os.path.join(T.GEN_DIR, name + ".hevc"), "rb").read()
    for seed in range(40):
        bad = T.corrupt(data, seed * 7919 + len(name))
        for fn in (hb.parse_picture, hb.parse_picture_k0):
            try:
                fn(bad, T._fmt(bad), host_only=True)
                ok += 1
            except hb.HeifCudaError:
                rejected += 1
print("parsed", ok, "rejected", rejected)
"""


_CONTAINER_CHILD = r"""
import sys, os, glob
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import numpy as np
import heif_b200 as hb
root = sys.argv[3]
files = sorted(glob.glob(os.path.join(root, "tests", "golden", "heic", "*.heic")))[::5] + sorted(glob.glob(os.path.join(root, "tests", "golden", "iovl", "*.heic")))[:2]
ok = rej = 0
for path in files:
    data = open(path, "rb").read()
    meta_end = data.find(b"mdat")
    meta_end = meta_end if meta_end > 0 else len(data)
    for seed in range(60):
        rng = np.random.default_rng(seed * 131 + len(data))
        d = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            i = int(rng.integers(0, meta_end))
            mode = int(rng.integers(0, 4))
            if mode == 0: d[i] = int(rng.integers(0, 256))
            elif mode == 1: d[i] = 0xff
            elif mode == 2: d[i] = 0
            else: d[i:i + 4] = int(rng.integers(0, 2 ** 32)).to_bytes(4, "big")
        if rng.integers(0, 5) == 0:
            d = d[:int(rng.integers(8, len(d)))]
        try:
            hf = hb.HeifFile(bytes(d), host_only=True)
            pid = hf.primary_id
            info = hf.image_info(pid)
            for t in (hf.grid_tiles(pid) if info.is_grid else [pid])[:4]:
                try: hf.coded_stream(t)
                except hb.HeifCudaError: pass
            try: hf.overlay(pid)
            except hb.HeifCudaError: pass
            if info.alpha_id:
                try: hf.image_info(info.alpha_id)
                except hb.HeifCudaError: pass
            hf.close()
            ok += 1
        except hb.HeifCudaError:
            rej += 1
print("read", ok, "rejected", rej)
"""


def test_container_reader_survives_damaged_files():
    """ISO-BMFF level: random bytes / sizes / truncations in the meta box of bundled files. The reader (heif_reader.cc)
    either answers or reports; a crash fails the child. tools/asan_fuzz.sh runs the same inputs (and the slice-data ones)
    against an AddressSanitizer + UBSan build of the host library."""
    r = subprocess.run([sys.executable, "-c", _CONTAINER_CHILD, os.path.join(ROOT, "heif-decoder-lib_b200"), os.path.join(ROOT, "tests"), ROOT],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    assert "read" in r.stdout and "rejected" in r.stdout


def test_cpu_parsers_survive_damaged_streams():
    r = subprocess.run([sys.executable, "-c", _CHILD, os.path.join(ROOT, "heif-decoder-lib_b200"), os.path.join(ROOT, "tests")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    assert "parsed" in r.stdout


@pytest.mark.gpu
def test_gpu_k0_damaged_pictures_do_not_disturb_their_batch():
    eng = hb.Engine(0)
    names = _victims()
    good, bad = [], []
    for name in names:
        data = open(os.path.join(GEN_DIR, name + ".hevc"), "rb").read()
        k = hb.K0Picture(data, _fmt(data))
        if k.eligible:
            good.append((name, k))
        for seed in range(12):
            try:
                kb = hb.K0Picture(corrupt(data, seed * 104729 + len(name)), _fmt(data))
            except hb.HeifCudaError:
                continue                                   # header damage: rejected on the host
            if kb.eligible:
                bad.append(kb)
    assert good and bad
    # 1. every damaged picture alone: decodes or is reported, the engine stays usable
    reported = 0
    for kb in bad:
        b = eng.batch()
        c = b.add_canvas(kb.pic.crop_w, kb.pic.crop_h, kb.pic.chroma_format, kb.pic.bit_depth_y)
        b.add_k0_picture(kb, c)
        b.upload()
        try:
            b.reconstruct(hb.STAGE_ALL)
            b.read_plane(c, 0)
        except hb.HeifCudaError as e:
            assert "device parser" in str(e) or "malformed" in str(e) or "capacity" in str(e), str(e)
            reported += 1
        b.close()
    # 2. the undamaged pictures afterwards, in one batch: still bit-exact
    b = eng.batch()
    items = []
    for name, k in good:
        c = b.add_canvas(k.pic.crop_w, k.pic.crop_h, k.pic.chroma_format, k.pic.bit_depth_y)
        b.add_k0_picture(k, c)
        items.append((name, c, k.pic))
    b.upload()
    b.reconstruct(hb.STAGE_ALL)
    for name, c, pic in items:
        planes = [b.read_plane(c, k) for k in range(3 if pic.chroma_format else 1)]
        dt = np.uint8 if pic.bit_depth_y == 8 else np.dtype("<u2")
        assert hashlib.md5(b"".join(p.astype(dt).tobytes() for p in planes)).hexdigest() == GEN_META[name]["yuv_md5"], name
    b.close()
    eng.close()
    print("damaged pictures:", len(bad), "reported:", reported)
