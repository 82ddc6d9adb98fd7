"""CPU tests: the product's host front-end (CABAC parse -> records) + the oracle restatement must
reproduce the reference's golden outputs bit-exactly. This is what pins the oracle (and the host
parser) before any CUDA result is compared against it."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import heif_b200 as hb
import oracle_lib
from conftest import ROOT, read_stream

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.mark.parametrize("name", sorted(GOLDEN["yuv_md5"]))
def test_annexb_streams_match_reference_yuv(name):
    rec = hb.parse_picture(read_stream(name), hb.STREAM_ANNEXB, host_only=True)
    planes, _ = oracle_lib.reconstruct(rec)
    assert md5(oracle_lib.planes_bytes(planes, rec.pic.bit_depth_y)) == GOLDEN["yuv_md5"][name]


def oracle_csc(planes, pic, matrix, full_range, out_format, alpha=None):
    O = oracle_lib.lib()
    h, w = planes[0].shape
    bpp = {0: 3, 1: 4, 2: 6, 3: 8, 4: 6, 5: 8}[out_format]
    out = np.zeros((h, w * bpp), np.uint8)
    pl = [np.ascontiguousarray(p, np.uint16) for p in planes]
    a = np.ascontiguousarray(alpha, np.uint16) if alpha is not None else None
    rc = O.hc_oracle_csc(C.c_void_p(pl[0].ctypes.data), C.c_void_p(pl[1].ctypes.data if len(pl) > 1 else None),
                         C.c_void_p(pl[2].ctypes.data if len(pl) > 2 else None), C.c_void_p(a.ctypes.data if a is not None else None),
                         pl[0].shape[1], pl[1].shape[1] if len(pl) > 1 else 0, a.shape[1] if a is not None else 0,
                         w, h, pic.chroma_format, pic.bit_depth_y, matrix, int(pic.colour_primaries), int(full_range), out_format,
                         C.c_void_p(out.ctypes.data), C.c_size_t(out.strides[0]))
    assert rc == 0
    return out


HEIC_CASES = [(f, item) for f in sorted(GOLDEN["heic"]) for item in sorted(GOLDEN["heic"][f])]


@pytest.mark.parametrize("fname,item", HEIC_CASES)
def test_heic_items_match_reference(fname, item):
    g = GOLDEN["heic"][fname][item]
    hf = hb.HeifFile(read_stream(fname), host_only=True)
    assert int(item) in hf.top_level_ids()
    rec = hb.parse_picture(hf.coded_stream(int(item)), host_only=True)
    pic = rec.pic
    assert (pic.crop_w, pic.crop_h) == (g["width"], g["height"])
    planes, _ = oracle_lib.reconstruct(rec)
    assert md5(oracle_lib.planes_bytes(planes, pic.bit_depth_y)) == g["planes_md5"]
    # colour conversion with the nclx the reference attaches to the decoded image (VUI, else 2/2/2 limited)
    rgb = oracle_csc(planes, pic, pic.matrix_coeffs, pic.full_range, 0)
    assert md5(rgb.tobytes()) == g["rgb_md5"]
    rgba = oracle_csc(planes, pic, pic.matrix_coeffs, pic.full_range, 1)
    assert md5(rgba.tobytes()) == g["rgba_md5"]


def test_container_reader_lists_items():
    hf = hb.HeifFile(read_stream("example.heic"), host_only=True)
    assert hf.top_level_ids() == [20004, 20006]
    assert hf.primary_id == 20004
    info = hf.image_info(20004)
    assert (info.width, info.height, info.is_grid) == (1280, 854, 0)


def test_malformed_inputs_fail_cleanly():
    with pytest.raises(hb.HeifCudaError):
        hb.HeifFile(b"\x00\x00\x00\x08junk", host_only=True)
    with pytest.raises(hb.HeifCudaError):
        hb.parse_picture(b"", host_only=True)
    data = read_stream("BasketballDrive_1920x1080_32.265")
    with pytest.raises(hb.HeifCudaError):
        hb.parse_picture(data[: len(data) // 2], hb.STREAM_ANNEXB, host_only=True)  # truncated slice data


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libheifref.so")),
                    reason="oracle/_ref (reference build) not present")
@pytest.mark.parametrize("fname,item", HEIC_CASES)
def test_heic_items_match_live_reference(fname, item):
    """Same check against the reference library itself (when oracle/_ref was built in this container)."""
    import refheif as R
    data = read_stream(fname)
    ref = R.decode(data, R.COLORSPACE_UNDEFINED, R.CHROMA_UNDEFINED, item_id=int(item))
    hf = hb.HeifFile(data, host_only=True)
    rec = hb.parse_picture(hf.coded_stream(int(item)), host_only=True)
    planes, _ = oracle_lib.reconstruct(rec)
    for p, k in zip(planes, ("Y", "Cb", "Cr")):
        assert p.astype(np.uint8).tobytes() == ref[k][0]


def test_oracle_bilinear_upsampling_reproduces_the_reference_vectors():
    """tests/conversion.cc:645-669 ("Bilinear upsampling"): the only colour-path op the reference pins with exact values —
    a 4x4 4:2:0 image, Cb / Cr 2x2 -> 4x4 through Op_YCbCr420_bilinear_to_YCbCr444."""
    O = oracle_lib.lib()
    cases = [([10, 40, 100, 240], [10, 18, 33, 40, 33, 47, 76, 90, 78, 106, 162, 190, 100, 135, 205, 240]),
             ([255, 200, 50, 0], [255, 241, 214, 200, 204, 190, 163, 150, 101, 88, 63, 50, 50, 38, 13, 0])]
    for src, want in cases:
        a = np.array(src, np.uint16).reshape(2, 2)
        out = np.zeros((4, 4), np.uint16)
        O.hc_oracle_bilinear_420(C.c_void_p(a.ctypes.data), 2, C.c_void_p(out.ctypes.data), 4, 4, 4)
        assert out.flatten().tolist() == want
