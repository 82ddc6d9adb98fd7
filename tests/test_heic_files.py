"""HEIC file fixtures (tests/golden/heic/, made by tests/golden/make_heic.py and decoded by the
unmodified reference's heif_decode_image): grids, limited-range tiles, colr override, 4:2:2/4:4:4,
10/12 bit, alpha auxiliary images -> interleaved RGB / RGBA / RRGGBB(AA)_LE.
CPU test: container reader + host parser + oracle reproduce the reference's MD5s.
GPU test: the native batch path (hc_heic_job) reproduces them too."""
import hashlib
import json
import os

import numpy as np
import pytest

import heif_b200 as hb
import heic_oracle
from conftest import ROOT

DIR = os.path.join(ROOT, "tests", "golden", "heic")
META = json.load(open(os.path.join(ROOT, "tests", "golden", "heic.json")))
NAMES = sorted(META)
FORMATS = {"rgb": hb.OUT_RGB, "rgba": hb.OUT_RGBA, "rrggbb_le": hb.OUT_RRGGBB_LE, "rrggbbaa_le": hb.OUT_RRGGBBAA_LE}
# every fixture x every interleaved chroma of heif_decode_image, the bit-depth changing ones included (10/12-bit images to
# RGB(A) 8 through Op_to_sdr_planes, 8-bit images to RRGGBB(AA) through Op_to_hdr_planes) and both byte orders
FMETA = json.load(open(os.path.join(ROOT, "tests", "golden", "heic_formats.json")))
ALL_FORMATS = dict(FORMATS, rrggbb_be=hb.OUT_RRGGBB_BE, rrggbbaa_be=hb.OUT_RRGGBBAA_BE)


def load(name):
    return open(os.path.join(DIR, name + ".heic"), "rb").read()


def md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference(name):
    m = META[name]
    planes, alpha, cf, bd, _ = heic_oracle.decode_planes(load(name))
    dt = np.uint8 if bd == 8 else np.dtype("<u2")
    allp = planes + ([alpha] if alpha is not None else [])
    assert md5(b"".join(p.astype(dt).tobytes() for p in allp)) == m["planes_md5"]
    for key, fmt in FORMATS.items():
        if key + "_md5" in m:
            assert md5(heic_oracle.decode_rgb(load(name), fmt).tobytes()) == m[key + "_md5"], key


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_in_every_output_format(name):
    data = load(name)
    checked = 0
    for key, fmt in ALL_FORMATS.items():
        if key + "_md5" in FMETA[name]:
            assert md5(heic_oracle.decode_rgb(data, fmt).tobytes()) == FMETA[name][key + "_md5"], key
            checked += 1
    assert checked >= 4, FMETA[name]


def _premultiply(rgba):
    """PREMULTI_PIXEL of heif_image_rgba_premultiply_alpha (pixelimage.cc:896-941): (v * A + 128) >> 8 on R, G, B"""
    px = rgba.reshape(rgba.shape[0], -1, 4).astype(np.uint32)
    px[:, :, :3] = (px[:, :, :3] * px[:, :, 3:4] + 128) >> 8
    return px.astype(np.uint8).reshape(rgba.shape)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_premultiplied_rgba_matches_reference(name):
    if "rgba_premultiplied_md5" not in FMETA[name]:
        pytest.skip(FMETA[name].get("rgba_premultiplied_error", "no golden"))
    assert md5(_premultiply(heic_oracle.decode_rgb(load(name), hb.OUT_RGBA)).tobytes()) == FMETA[name]["rgba_premultiplied_md5"]


def test_prem_reference_is_reported():
    hf = hb.HeifFile(load("alpha_prem_420_8"), host_only=True)
    assert hf.image_info(hf.primary_id).premultiplied_alpha == 1
    hf = hb.HeifFile(load("alpha_420_8"), host_only=True)
    assert hf.image_info(hf.primary_id).premultiplied_alpha == 0


BMETA = json.load(open(os.path.join(ROOT, "tests", "golden", "heic_formats_bilinear.json")))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_with_bilinear_upsampling(name):
    """heif_decode_image with the colour conversion options of heif-dec -C bilinear (bilinear chroma upsampling, only the
    preferred algorithm): every fixture, every interleaved format the reference converts."""
    data = load(name)
    cf = heic_oracle.decode_planes(data)[2]
    for key, fmt in ALL_FORMATS.items():
        if key + "_md5" in BMETA[name] and not (cf == 0 and fmt > hb.OUT_RGBA):
            assert md5(heic_oracle.decode_rgb(data, fmt, upsampling=1).tobytes()) == BMETA[name][key + "_md5"], key


def _iovl_file(chroma_format):
    from tools import heif_writer as W, hevcenc
    b = W.HeifBuilder()
    kids = [b.add_hevc_image(hevcenc.encode(hevcenc.synth_image(64, 64, chroma_format, 8, 70 + k), chroma_format=chroma_format, bit_depth=8, seed=70 + k),
                             64, 64, chroma_format, 8, hidden=True) for k in range(2)]
    b.primary = b.add_overlay(kids, 100, 80, [(3, 5), (30, 12)], background=(0x1234, 0x8000, 0xffff, 0xffff))
    return b.serialize()


def test_iovl_items_are_reported_not_decoded():
    """An 'iovl' derived image is outside the path: the reader names the item type instead of
    guessing, and the unmodified reference itself only composes overlays whose children are 4:4:4."""
    data = _iovl_file(1)
    hf = hb.HeifFile(data, host_only=True)
    with pytest.raises(hb.HeifCudaError, match="not an HEVC image"):
        hf.coded_stream(hf.primary_id)
    import refheif as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    with pytest.raises(RuntimeError, match="Unsupported color conversion"):
        R.decode(data, R.COLORSPACE_RGB, R.CHROMA_RGB)
    assert R.decode(_iovl_file(3), R.COLORSPACE_RGB, R.CHROMA_RGB)["interleaved"][1:] == (100, 80)


@pytest.fixture(scope="module")
def engine():
    e = hb.Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("device_parse", [1, 0])
@pytest.mark.parametrize("want_alpha", [False, True])
def test_gpu_heic_job_matches_reference(engine, want_alpha, device_parse):
    """All fixture files in ONE job (one upload, one launch sequence for every tile of every file); slice data
    parsed by K0 on the GPU (device_parse=1, the default) or by the host CABAC parser (0)."""
    engine.set_option("device_parse", device_parse)
    engine.set_option("host_share_pct", 0)     # not "auto": a job this small would otherwise pick the host parser by itself
    job = hb.HeicJob(engine, [load(n) for n in NAMES], want_alpha=want_alpha, threads=4)
    engine.set_option("device_parse", 1)
    engine.set_option("host_share_pct", -1)
    job.upload()
    job.run()
    for i, name in enumerate(NAMES):
        m, d = META[name], job.descs[i]
        assert (d.width, d.height, d.bit_depth, bool(d.has_alpha)) == (m["width"], m["height"], m["bit_depth"], m["has_alpha"])
        key = {hb.OUT_RGB: "rgb", hb.OUT_RGBA: "rgba", hb.OUT_RRGGBB_LE: "rrggbb_le", hb.OUT_RRGGBBAA_LE: "rrggbbaa_le"}[d.out_format]
        assert md5(job.read_rgb(i).tobytes()) == m[key + "_md5"], (name, key)
        dt = np.uint8 if d.bit_depth == 8 else np.dtype("<u2")
        planes = [job.read_plane(i, k) for k in range(3 if d.chroma_format else 1)]
        if d.has_alpha:
            planes.append(job.read_plane(i, 3))
        assert md5(b"".join(p.astype(dt).tobytes() for p in planes)) == m["planes_md5"], name
    assert job.launch_count >= 6
    job.close()


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(ALL_FORMATS))
def test_gpu_heic_job_every_output_format(engine, key):
    """hc_heic_job with HC_OUTPUT_FORMAT: all fixtures to ONE interleaved format whatever their bit depth — K5 with the
    reference's Op_to_sdr_planes / Op_to_hdr_planes fused before or after the matrix (csc_select.cc)."""
    names = [n for n in NAMES if key + "_md5" in FMETA[n]]
    assert len(names) >= 40
    job = hb.HeicJob(engine, [load(n) for n in names], threads=4, out_format=ALL_FORMATS[key])
    job.upload()
    job.run()
    bad = [n for i, n in enumerate(names) if md5(job.read_rgb(i).tobytes()) != FMETA[n][key + "_md5"]]
    job.close()
    assert not bad, (key, bad)


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["rgb", "rgba", "rrggbb_le", "rrggbbaa_be"])
def test_gpu_heic_job_bilinear_chroma_upsampling(engine, key):
    """engine option chroma_upsampling = HC_UPSAMPLE_BILINEAR: K5 fetches chroma through the reference's bilinear ops
    (border rules included), bit-exact against the reference run with the same option."""
    fmt = ALL_FORMATS[key]
    names = [n for n in NAMES if key + "_md5" in BMETA[n] and not (META[n].get("chroma_mono") or n.startswith("single_mono") and fmt > hb.OUT_RGBA)]
    engine.set_option("chroma_upsampling", 1)
    try:
        job = hb.HeicJob(engine, [load(n) for n in names], threads=4, out_format=fmt)
    finally:
        engine.set_option("chroma_upsampling", 0)
    job.upload()
    job.run()
    bad = [n for i, n in enumerate(names) if md5(job.read_rgb(i).tobytes()) != BMETA[n][key + "_md5"]]
    job.close()
    assert len(names) >= 35 and not bad, (key, bad)


@pytest.mark.gpu
def test_gpu_heic_job_premultiplied_alpha(engine):
    """engine option premultiply_alpha: RGBA output multiplied by its alpha in K5, bit-exact with the reference's
    heif_image_rgba_premultiply_alpha on its RGBA decode; a file that is already premultiplied ('prem') is left alone and
    reported as such."""
    names = [n for n in NAMES if "rgba_premultiplied_md5" in FMETA[n]]
    engine.set_option("premultiply_alpha", 1)
    try:
        job = hb.HeicJob(engine, [load(n) for n in names], threads=4, out_format=hb.OUT_RGBA)
    finally:
        engine.set_option("premultiply_alpha", 0)
    job.upload()
    job.run()
    bad = []
    for i, n in enumerate(names):
        assert job.descs[i].premultiplied_alpha == 1, n
        want = FMETA[n]["rgba_md5"] if n == "alpha_prem_420_8" else FMETA[n]["rgba_premultiplied_md5"]
        if md5(job.read_rgb(i).tobytes()) != want:
            bad.append(n)
    job.close()
    assert not bad, bad
    job = hb.HeicJob(engine, [load("alpha_prem_420_8"), load("alpha_420_8")], threads=2, out_format=hb.OUT_RGBA)
    assert [d.premultiplied_alpha for d in job.descs] == [1, 0]
    job.close()


@pytest.mark.gpu
def test_gpu_decode_heic_python_path(engine):
    for name in ("grid_300x200_t128", "single_420_8_novui", "alpha_420_8"):
        rgb = hb.decode_heic(engine, load(name), hb.OUT_RGB)
        assert md5(rgb.tobytes()) == META[name]["rgb_md5"], name


@pytest.mark.gpu
@pytest.mark.parametrize("host_share", [-1, 0, 50, 100])
def test_gpu_decode_stream_pipelined_batches(engine, host_share):
    """hc_heic_decode_stream: batches of 4 files, submit / deliver pipelined three batches deep, slice data parsed
    by K0 with the host threads taking `host_share` per cent of the coded items meanwhile (-1: automatic);
    every image arrives once, in order, bit-exact."""
    engine.set_option("host_share_pct", host_share)
    files = [load(n) for n in NAMES] * 2
    seen = []

    def on_image(index, desc, rows):
        name = NAMES[index % len(NAMES)]
        key = {hb.OUT_RGB: "rgb", hb.OUT_RGBA: "rgba", hb.OUT_RRGGBB_LE: "rrggbb_le", hb.OUT_RRGGBBAA_LE: "rrggbbaa_le"}[desc.out_format]
        seen.append((index, md5(rows.tobytes()) == META[name][key + "_md5"]))

    st = hb.decode_stream(engine, files, on_image, want_alpha=False, threads=4, files_per_batch=4)
    engine.set_option("host_share_pct", -1)
    assert [i for i, _ in seen] == list(range(len(files)))
    assert all(ok for _, ok in seen)
    assert st["batches"] == (len(files) + 3) // 4 and st["pixels"] == 2 * sum(META[n]["width"] * META[n]["height"] for n in NAMES)
    assert st["launches"] > 0 and st["bytes_d2h"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("nbatches", [1, 2, 3, 4])
def test_gpu_decode_stream_fill_and_drain(engine, nbatches):
    """Streams shorter than, as long as and longer than the pipeline depth (3): every image once, in order, bit-exact."""
    names = (NAMES * 2)[:3 * nbatches - 1]          # the last batch is ragged
    seen = []

    def on_image(index, desc, rows):
        key = {hb.OUT_RGB: "rgb", hb.OUT_RGBA: "rgba", hb.OUT_RRGGBB_LE: "rrggbb_le", hb.OUT_RRGGBBAA_LE: "rrggbbaa_le"}[desc.out_format]
        seen.append((index, md5(rows.tobytes()) == META[names[index]][key + "_md5"]))

    st = hb.decode_stream(engine, [load(n) for n in names], on_image, want_alpha=False, threads=2, files_per_batch=3)
    assert [i for i, _ in seen] == list(range(len(names))) and all(ok for _, ok in seen)
    assert st["batches"] == nbatches


@pytest.mark.gpu
def test_gpu_decode_stream_automatic_batch_size(engine):
    """files_per_batch = 0: the library picks the batch size from the size of the first image (about 800 MP per batch;
    these small fixtures all fit one batch of at most 512 files); every image once, in order, bit-exact."""
    names = NAMES * 3
    seen = []

    def on_image(index, desc, rows):
        key = {hb.OUT_RGB: "rgb", hb.OUT_RGBA: "rgba", hb.OUT_RRGGBB_LE: "rrggbb_le", hb.OUT_RRGGBBAA_LE: "rrggbbaa_le"}[desc.out_format]
        seen.append((index, md5(rows.tobytes()) == META[names[index]][key + "_md5"]))

    st = hb.decode_stream(engine, [load(n) for n in names], on_image, want_alpha=False, threads=4, files_per_batch=0)
    assert [i for i, _ in seen] == list(range(len(names))) and all(ok for _, ok in seen)
    assert st["batches"] == 1


@pytest.mark.gpu
def test_gpu_decode_stream_external_destinations(engine):
    """hc_heic_decode_stream_ext: the reference's ext_dst semantics (heif.h:1605-1615, pixelimage.cc:221-266) — final pixels in
    the caller's buffer with the caller's stride when it is large enough, the library's own memory otherwise."""
    names = NAMES[:9]
    files = [load(n) for n in names]
    dests, kinds = [], []
    for k, n in enumerate(names):
        m = META[n]
        bpp = 3 if m["bit_depth"] == 8 else 6
        row = m["width"] * bpp
        if k % 3 == 0:
            dests.append(np.zeros((m["height"], row + 40), np.uint8)); kinds.append("padded")      # larger stride
        elif k % 3 == 1:
            dests.append(np.zeros((m["height"] - 1, row), np.uint8)); kinds.append("small")         # too small: ignored
        else:
            dests.append(None); kinds.append("none")
    seen = {}

    def on_image(index, desc, rows):
        key = "rgb_md5" if desc.out_format == hb.OUT_RGB else "rrggbb_le_md5"
        row = desc.width * desc.bytes_per_pixel
        ext = dests[index] is not None and rows.ctypes.data == dests[index].ctypes.data
        seen[index] = (ext, rows.shape[1], md5(np.ascontiguousarray(rows[:, :row]).tobytes()) == META[names[index]][key])

    hb.decode_stream(engine, files, on_image, threads=2, files_per_batch=4, dests=dests)
    for k, kind in enumerate(kinds):
        ext, stride, ok = seen[k]
        assert ok, names[k]
        assert ext == (kind == "padded"), (names[k], kind)
        if kind == "padded":
            m = META[names[k]]
            key = "rgb_md5" if m["bit_depth"] == 8 else "rrggbb_le_md5"
            row = m["width"] * (3 if m["bit_depth"] == 8 else 6)
            assert stride == row + 40 and md5(np.ascontiguousarray(dests[k][:, :row]).tobytes()) == m[key]


@pytest.mark.gpu
def test_gpu_decode_stream_reports_bad_file(engine):
    files = [load(NAMES[0]), b"not a heic file at all", load(NAMES[1])]
    with pytest.raises(hb.HeifCudaError):
        hb.decode_stream(engine, files, None, files_per_batch=1)


def _heic_with_rejected_slice_data():
    """A well-formed HEIC whose slice data the K0 parser rejects (checked with the CPU build of the same parser), and the
    undamaged file: (bad, good)."""
    import test_fuzz as T
    from tools import heif_writer as W
    name = "wpp"
    meta = T.GEN_META[name]
    data = open(os.path.join(T.GEN_DIR, name + ".hevc"), "rb").read()
    for seed in range(200):
        bad = T.corrupt(data, seed * 7919 + 11)
        if len(bad) != len(data):
            continue                      # truncations change the parameter-set scan of the writer; keep the length
        try:
            hb.parse_picture_k0(bad, T._fmt(bad), host_only=True)
        except hb.HeifCudaError:
            return (W.single_image(bad, meta["width"], meta["height"], 1, 8), W.single_image(data, meta["width"], meta["height"], 1, 8))
    raise AssertionError("no corruption of %s is rejected by the parser" % name)


@pytest.mark.gpu
def test_gpu_decode_stream_isolates_bad_files(engine):
    """hc_heic_decode_stream_ext with file_status: a file that is not a HEIC at all (fails in the container parse on the
    host) and a file whose slice data the DEVICE parser rejects each get an error code; every other file of their batches
    is delivered bit-exactly, in every position of the batch."""
    bad_slices, good_twin = _heic_with_rejected_slice_data()
    files = [load(NAMES[0]), b"not a heic file at all", load(NAMES[1]), bad_slices, good_twin, load(NAMES[2]), bad_slices]
    expect_bad = {1, 3, 6}
    got = {}

    def on_image(index, desc, rows):
        got[index] = hashlib.md5(np.ascontiguousarray(rows[:, :desc.width * desc.bytes_per_pixel]).tobytes()).hexdigest()

    for per_batch in (3, 7, 1):
        got.clear()
        st = hb.decode_stream(engine, files, on_image, files_per_batch=per_batch, isolate_errors=True)
        assert {k for k, v in enumerate(st["file_status"]) if v != 0} == expect_bad, (per_batch, st["file_status"])
        assert st["files_failed"] == 3 and st["first_error"]
        assert set(got) == set(range(len(files))) - expect_bad
        want = {}
        for k in got:   # each good file alone through the same API
            hb.decode_stream(engine, [files[k]], lambda i, d, r, k=k: want.__setitem__(
                k, hashlib.md5(np.ascontiguousarray(r[:, :d.width * d.bytes_per_pixel]).tobytes()).hexdigest()), files_per_batch=1)
        assert got == want, per_batch
    # without file_status the call still fails as a whole
    with pytest.raises(hb.HeifCudaError):
        hb.decode_stream(engine, files, None, files_per_batch=3)


def test_rejected_slice_data_fixture_is_a_wellformed_container():
    bad, good = _heic_with_rejected_slice_data()
    for data in (bad, good):
        hf = hb.HeifFile(data, host_only=True)
        assert hf.image_info(hf.primary_id).width == 264
        hf.close()
