"""Parameter-set bookkeeping of the host front-end (csrc/host/hevc_parse.cc push_nal), the cases a crafted HEIC can reach
through hvcC + item data: a re-sent SPS invalidates the PPSs that refer to it (libde265 does the same,
decctx.cc:575-584), parameter sets may not arrive inside a picture, and a monochrome SPS's chroma bit depth is ignored."""
import os

import pytest

import heif_b200 as hb
from conftest import ROOT

GEN_DIR = os.path.join(ROOT, "tests", "golden", "generated")


def nals(data):
    """Annex-B or 4-byte-length-prefixed stream -> list of NAL unit payloads."""
    out, i, n = [], 0, len(data)
    if not (data[:3] == b"\x00\x00\x01" or data[:4] == b"\x00\x00\x00\x01"):
        while i + 4 <= n:
            ln = int.from_bytes(data[i:i + 4], "big")
            out.append(bytes(data[i + 4:i + 4 + ln]))
            i += 4 + ln
        return out
    starts = []
    while i + 3 <= n:
        if data[i] == 0 and data[i + 1] == 0 and data[i + 2] == 1:
            starts.append(i + 3)
            i += 3
        else:
            i += 1
    for k, s in enumerate(starts):
        e = starts[k + 1] - 3 if k + 1 < len(starts) else n
        while e > s and data[e - 1] == 0:
            e -= 1
        out.append(bytes(data[s:e]))
    return out


def annexb(units):
    return b"".join(b"\x00\x00\x00\x01" + u for u in units)


def nal_type(u):
    return (u[0] >> 1) & 0x3F


def split(name):
    u = nals(open(os.path.join(GEN_DIR, name + ".hevc"), "rb").read())
    sps = [x for x in u if nal_type(x) == 33]
    pps = [x for x in u if nal_type(x) == 34]
    vcl = [x for x in u if nal_type(x) < 32]
    assert sps and pps and vcl
    return sps, pps, vcl


def test_resent_sps_invalidates_its_pps():
    """SPS(small), PPS, SPS(same id, larger picture), slice: the stale PPS must not be paired with the new SPS (its derived
    tables are sized for the old one — heap overflow before the fix); the slice now refers to a missing PPS."""
    sps_a, pps_a, vcl_a = split("ctb16_sao")
    sps_b, _, _ = split("base_420_8")
    bad = annexb(sps_a + pps_a + sps_b + vcl_a)
    with pytest.raises(hb.HeifCudaError) as ei:
        hb.parse_picture(bad, hb.STREAM_ANNEXB, host_only=True)
    assert "missing PPS" in str(ei.value)
    # the K0 preparation walks the same code
    with pytest.raises(hb.HeifCudaError):
        hb.K0Picture(bad, hb.STREAM_ANNEXB, host_only=True)
    # re-sending SPS and PPS together (what encoders do before every IRAP picture) keeps decoding
    ok = annexb(sps_a + pps_a + sps_a + pps_a + vcl_a)
    want = hb.parse_picture(annexb(sps_a + pps_a + vcl_a), hb.STREAM_ANNEXB, host_only=True)
    got = hb.parse_picture(ok, hb.STREAM_ANNEXB, host_only=True)
    assert got.pic.blk_count == want.pic.blk_count and got.pic.coeff_count == want.pic.coeff_count


def test_parameter_sets_inside_a_picture_are_rejected():
    sps, pps, vcl = split("slices_dep")
    assert len(vcl) > 1
    for extra, what in ((sps, "SPS"), (pps, "PPS")):
        bad = annexb(sps + pps + vcl[:1] + extra + vcl[1:])
        with pytest.raises(hb.HeifCudaError) as ei:
            hb.parse_picture(bad, hb.STREAM_ANNEXB, host_only=True)
        assert what + " inside a picture" in str(ei.value)


def test_monochrome_ignores_the_coded_chroma_bit_depth():
    """A monochrome SPS still codes bit_depth_chroma_minus8; the picture type must follow luma alone (a
    4:0:0 8-bit picture with chroma depth 12 selected 16-bit kernels for an 8-bit canvas)."""
    for name, depth in (("mono_8", 8), ("mono_12", 12)):
        rec = hb.parse_picture(annexb(nals(open(os.path.join(GEN_DIR, name + ".hevc"), "rb").read())), hb.STREAM_ANNEXB, host_only=True)
        assert rec.pic.chroma_format == 0 and rec.pic.bit_depth_c == rec.pic.bit_depth_y == depth


def test_a_reused_parser_starts_every_picture_from_clean_cu_state():
    """One parser object decodes many pictures (one per host thread in hc_heic_job). CuQpDeltaVal is read by every QP
    derivation but only written when the stream codes cu_qp_delta: after example.heic (cu_qp_delta on) a picture without it
    came out with every QP shifted. The records of the second picture must equal those of a fresh parser."""
    import ctypes as C
    import numpy as np
    from heif_b200 import _lib
    from heif_b200.api import Records
    from conftest import STREAMS
    L = _lib.load(True)

    def primary_stream(path):
        hf = hb.HeifFile(open(path, "rb").read(), host_only=True)
        return hf.coded_stream(hf.primary_id)

    first = primary_stream(os.path.join(STREAMS, "example.heic"))
    second = primary_stream(os.path.join(STREAMS, "test_832x480.heic"))
    fresh = hb.parse_picture(second, host_only=True)
    p = L.hc_parser_new()
    try:
        assert L.hc_parser_push(p, first, len(first), hb.STREAM_LENGTH_PREFIXED) == 0
        Records(L, L.hc_parser_take_picture(p))
        assert L.hc_parser_push(p, second, len(second), hb.STREAM_LENGTH_PREFIXED) == 0
        reused = Records(L, L.hc_parser_take_picture(p))
    finally:
        L.hc_parser_free(p)
    for name, size in (("ctus", 44), ("blks", 16), ("tbs", 16), ("coeffs", 4), ("edge_map", 1), ("qp_map", 1)):
        pa, na = fresh.array(name)
        pb, nb = reused.array(name)
        assert na == nb, name
        xa = np.ctypeslib.as_array(C.cast(pa, C.POINTER(C.c_uint8)), shape=(na * size,))
        xb = np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_uint8)), shape=(nb * size,))
        assert np.array_equal(xa, xb), name
