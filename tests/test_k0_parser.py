"""K0 — the DEVICE CABAC parser (csrc/kernels/k0_core.cuh) — validated on the CPU: the same core statements that run
one-warp-per-substream on the GPU are executed sequentially by hc_parse_picture_k0 and their records must equal the
host parser's records array by array (blocks, transform blocks, coefficients, CTU records, edge / QP maps), for every
stream fixture K0 accepts; the others must be declined with a reason (they stay with the host parser)."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest

import heif_b200 as hb
from conftest import ROOT, STREAMS

GEN_DIR = os.path.join(ROOT, "tests", "golden", "generated")
HEIC_DIR = os.path.join(ROOT, "tests", "golden", "heic")
GEN = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GEN_DIR, "*.hevc")))
# fixtures using coding tools K0 leaves to the host parser (k0_core.cuh header)
DECLINED = {"bypass", "bypass_10", "pcm", "pcm_bypass_422", "tiles_2x2", "tiles_3x2_nolf", "slices_nolf_422_t"}

ELEM = {"ctus": 44, "blks": 16, "tbs": 16, "coeffs": 4, "edge_map": 1, "qp_map": 1}


def arrays(rec):
    out = {}
    for name, size in ELEM.items():
        p, n = rec.array(name)
        out[name] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n * size,)).copy() if n else np.zeros(0, np.uint8)
    return out


def check_equal(data, fmt):
    a = hb.parse_picture(data, fmt, host_only=True)
    b = hb.parse_picture_k0(data, fmt, host_only=True)
    pa, pb = a.pic, b.pic
    for f in ("width", "height", "crop_w", "crop_h", "chroma_format", "bit_depth_y", "log2_ctb", "ctbs_w", "ctbs_h", "blk_count",
              "tb_count", "coeff_count", "resid_count", "matrix_coeffs", "full_range"):
        assert getattr(pa, f) == getattr(pb, f), f
    assert (pa.flags | 0x000c) == (pb.flags | 0x000c)          # HAS_DEBLOCK / HAS_SAO are set conservatively by K0
    xa, xb = arrays(a), arrays(b)
    for name in ELEM:
        assert xa[name].shape == xb[name].shape, name
        if name == "tbs":                                         # hc_tb::pic is batch placement, not parse output
            ta, tb = xa[name].reshape(-1, 16).copy(), xb[name].reshape(-1, 16).copy()
            ta[:, 10:12] = 0
            tb[:, 10:12] = 0
            assert np.array_equal(ta, tb), name
        else:
            assert np.array_equal(xa[name], xb[name]), name


@pytest.mark.parametrize("name", [n for n in GEN if n not in DECLINED])
def test_k0_records_equal_host_parser_generated(name):
    check_equal(open(os.path.join(GEN_DIR, name + ".hevc"), "rb").read(), hb.STREAM_LENGTH_PREFIXED if False else _fmt(name))


def _fmt(name):
    data = open(os.path.join(GEN_DIR, name + ".hevc"), "rb").read()
    return hb.STREAM_ANNEXB if data[:3] == b"\x00\x00\x01" or data[:4] == b"\x00\x00\x00\x01" else hb.STREAM_LENGTH_PREFIXED


@pytest.mark.parametrize("name", sorted(DECLINED & set(GEN)))
def test_k0_declines_what_it_does_not_parse(name):
    with pytest.raises(hb.HeifCudaError) as ei:
        hb.parse_picture_k0(open(os.path.join(GEN_DIR, name + ".hevc"), "rb").read(), _fmt(name), host_only=True)
    assert "not eligible" in str(ei.value)


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(STREAMS, "*.265"))))
def test_k0_records_equal_host_parser_reference_streams(name):
    """the reference's own 1080p test streams (WPP, transform skip, TU depth 3)"""
    check_equal(open(os.path.join(STREAMS, name), "rb").read(), hb.STREAM_ANNEXB)


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(STREAMS, "*.heic")) +
                                        glob.glob(os.path.join(HEIC_DIR, "*.heic"))))
def test_k0_records_equal_host_parser_heic_items(name):
    path = os.path.join(STREAMS, name) if os.path.exists(os.path.join(STREAMS, name)) else os.path.join(HEIC_DIR, name)
    hf = hb.HeifFile(open(path, "rb").read(), host_only=True)
    info = hf.image_info(hf.primary_id)
    ids = hf.grid_tiles(hf.primary_id) if info.is_grid else [hf.primary_id]
    if info.alpha_id:
        ainfo = hf.image_info(info.alpha_id)
        ids = ids[:5] + (hf.grid_tiles(info.alpha_id)[:2] if ainfo.is_grid else [info.alpha_id])
    for i in ids[:7]:
        check_equal(hf.coded_stream(i), hb.STREAM_LENGTH_PREFIXED)


# ---------------------------------------------------------------------------------------------------------------
# GPU: the same core as kernel K0 (one CTA per substream chain); reconstruction of device-parsed pictures must equal
# the golden MD5s of the reference
import hashlib  # noqa: E402

GEN_META = json.load(open(os.path.join(ROOT, "tests", "golden", "generated.json")))


@pytest.fixture(scope="module")
def engine():
    e = hb.Engine(0)
    yield e
    e.close()


def _planes_md5(b, canvas, pic):
    planes = [b.read_plane(canvas, k) for k in range(3 if pic.chroma_format else 1)]
    dt = np.uint8 if pic.bit_depth_y == 8 else np.dtype("<u2")
    return hashlib.md5(b"".join(p.astype(dt).tobytes() for p in planes)).hexdigest()


@pytest.mark.gpu
def test_gpu_k0_all_eligible_generated_streams_in_one_batch(engine):
    names = [n for n in GEN if n not in DECLINED]
    b = engine.batch()
    items = []
    for n in names:
        k = hb.K0Picture(open(os.path.join(GEN_DIR, n + ".hevc"), "rb").read(), _fmt(n))
        assert k.eligible, (n, k.why_not)
        p = k.pic
        c = b.add_canvas(p.crop_w, p.crop_h, p.chroma_format, p.bit_depth_y)
        b.add_k0_picture(k, c)
        items.append((n, c, k))
    b.upload()
    b.reconstruct(hb.STAGE_ALL)
    assert b.stage_ms()["k0_parse"] > 0
    bad = [n for n, c, k in items if _planes_md5(b, c, k.pic) != GEN_META[n]["yuv_md5"]]
    b.close()
    assert not bad, bad


@pytest.mark.gpu
def test_gpu_k0_reference_1080p_streams_and_mixed_batch(engine):
    """device-parsed (K0) and host-parsed pictures side by side in one batch"""
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    b = engine.batch()
    items = []
    for i, name in enumerate(sorted(os.path.basename(p) for p in glob.glob(os.path.join(STREAMS, "*.265")))):
        data = open(os.path.join(STREAMS, name), "rb").read()
        if i % 2 == 0:
            k = hb.K0Picture(data, hb.STREAM_ANNEXB)
            assert k.eligible, k.why_not
            c = b.add_canvas(k.pic.crop_w, k.pic.crop_h, 1, 8)
            b.add_k0_picture(k, c)
            items.append((name, c, k.pic, k))
        else:
            r = hb.parse_picture(data, hb.STREAM_ANNEXB)
            c = b.add_canvas(r.pic.crop_w, r.pic.crop_h, 1, 8)
            b.add_picture(r, c)
            items.append((name, c, r.pic, r))
    b.upload()
    b.reconstruct(hb.STAGE_ALL)
    for name, c, pic, _ in items:
        assert _planes_md5(b, c, pic) == golden["yuv_md5"][name], name
    b.close()


@pytest.mark.gpu
def test_gpu_k0_malformed_slice_data_is_reported(engine):
    data = bytearray(open(os.path.join(GEN_DIR, "base_420_8.hevc"), "rb").read())
    for i in range(len(data) - 400, len(data) - 100):
        data[i] = 0xFF if data[i] != 0 else 0x80      # avoid creating start codes / emulation patterns
    try:
        k = hb.K0Picture(bytes(data), _fmt("base_420_8"))
    except hb.HeifCudaError:
        return                                           # header-level damage: rejected on the host already
    b = engine.batch()
    c = b.add_canvas(k.pic.crop_w, k.pic.crop_h, k.pic.chroma_format, k.pic.bit_depth_y)
    b.add_k0_picture(k, c)
    b.upload()
    try:
        b.reconstruct(hb.STAGE_ALL)      # either the damage still decodes (CABAC is forgiving) or it is reported
    except hb.HeifCudaError as e:
        assert "device parser" in str(e) or "malformed" in str(e)
    b.close()
