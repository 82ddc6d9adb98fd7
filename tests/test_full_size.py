"""BASELINE.json configurations at (or near) full size on the GPU, generated on the fly with the repo's own
content tools (tools/hevc_enc + tools/heif_writer; every stream is decodable by the unmodified reference):

  C2  4032x3024 grid of 512x512 tiles, 8-bit 4:2:0 -> RGB24           bit-exact vs the reference (or the oracle)
  C3  2048x1536 single 4:2:2 10-bit and 12-bit + alpha -> RRGGBBAA_LE  bit-exact vs the reference (or the oracle)
  C4  a list of 1920x1080 files through the pipelined stream API       every image equals the single-file decode
plus size-independent properties: decoding the same file twice in one batch gives identical pixels (idempotence /
no cross-picture interference), and a grid equals its tiles decoded alone and pasted (tile independence)."""
import hashlib
import os
import sys

import numpy as np
import pytest

import heif_b200 as hb
import heic_oracle
import refheif as R
from conftest import ROOT

sys.path.insert(0, ROOT)
from tools import heif_writer, hevcenc  # noqa: E402

pytestmark = pytest.mark.gpu


def md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.fixture(scope="module")
def engine():
    e = hb.Engine(0)
    yield e
    e.close()


def reference_rgb(data, out_format):
    """The unmodified reference when its build travelled with the repo (oracle/_ref), else the pinned oracle."""
    if R.available():
        chroma = {hb.OUT_RGB: R.CHROMA_RGB, hb.OUT_RGBA: R.CHROMA_RGBA, hb.OUT_RRGGBB_LE: R.CHROMA_RRGGBB_LE,
                  hb.OUT_RRGGBBAA_LE: R.CHROMA_RRGGBBAA_LE}[out_format]
        return R.decode(data, R.COLORSPACE_RGB, chroma)["interleaved"][0]
    return heic_oracle.decode_rgb(data, out_format).tobytes()


def test_c2_12mp_grid_rgb(engine):
    data = heif_writer.synth_grid_heic(4032, 3024, tile=512, seed=7, qp=26, wpp=1, sao=1, log2_ctb=6)
    job = hb.HeicJob(engine, [data, data], want_alpha=False, threads=0)
    job.upload()
    job.run()
    d = job.descs[0]
    assert (d.width, d.height, d.coded_pictures, d.out_format) == (4032, 3024, 48, hb.OUT_RGB)
    a, b = job.read_rgb(0), job.read_rgb(1)
    job.close()
    assert np.array_equal(a, b)                                    # same file twice in one batch
    assert md5(a.tobytes()) == md5(reference_rgb(data, hb.OUT_RGB))


@pytest.mark.parametrize("bit_depth", [10, 12])
def test_c3_422_hdr_with_alpha_rrggbbaa(engine, bit_depth):
    w, h = 2048, 1536
    planes = hevcenc.synth_image(w, h, 2, bit_depth, seed=11)
    colour = hevcenc.encode(planes, chroma_format=2, bit_depth=bit_depth, qp=24, wpp=1, sao=1, log2_ctb=6, seed=3)
    alpha = hevcenc.encode([hevcenc.synth_image(w, h, 0, bit_depth, seed=12)[0]], chroma_format=0, bit_depth=bit_depth, qp=30, wpp=1,
                           log2_ctb=6, seed=4)
    data = heif_writer.single_image(colour, w, h, 2, bit_depth, alpha_stream=alpha)
    job = hb.HeicJob(engine, [data], want_alpha=True, threads=0)
    job.upload()
    job.run()
    d = job.descs[0]
    assert (d.width, d.height, d.bit_depth, bool(d.has_alpha), d.out_format, d.bytes_per_pixel) == (w, h, bit_depth, True,
                                                                                                   hb.OUT_RRGGBBAA_LE, 8)
    got = job.read_rgb(0)
    job.close()
    assert md5(got.tobytes()) == md5(reference_rgb(data, hb.OUT_RRGGBBAA_LE))


def test_c4_stream_of_1080p_files(engine):
    files = []
    for seed in range(3):
        planes = hevcenc.synth_image(1920, 1080, 1, 8, seed=20 + seed)
        s = hevcenc.encode(planes, chroma_format=1, bit_depth=8, qp=28 + seed, wpp=1, sao=1, log2_ctb=6, seed=seed)
        files.append(heif_writer.single_image(s, 1920, 1080, 1, 8))
    want = [md5(reference_rgb(f, hb.OUT_RGB)) for f in files]
    order = [0, 1, 2, 2, 1, 0, 1, 0, 2, 0]
    seen = {}

    def on_image(index, desc, rows):
        assert (desc.width, desc.height) == (1920, 1080)
        seen[index] = md5(rows.tobytes())

    st = hb.decode_stream(engine, [files[k] for k in order], on_image, threads=0, files_per_batch=4)
    assert [seen[i] for i in range(len(order))] == [want[k] for k in order]
    assert st["batches"] == 3 and st["pixels"] == len(order) * 1920 * 1080


def test_grid_equals_independently_decoded_tiles(engine):
    """HEIF grid tiles are independent HEVC pictures (context.cc:2407-2415): the grid canvas must equal every
    tile decoded on its own and pasted at its offset."""
    data = heif_writer.synth_grid_heic(1024, 768, tile=256, seed=5, qp=30, wpp=0, sao=1, log2_ctb=5)
    job = hb.HeicJob(engine, [data], threads=2)
    job.upload()
    job.run()
    grid_y = job.read_plane(0, 0)
    job.close()
    hf = hb.HeifFile(data)
    tiles = hf.grid_tiles(hf.primary_id)
    for k, t in enumerate(tiles):
        rec = hb.parse_picture(hf.coded_stream(t))
        b = engine.batch()
        c = b.add_canvas(256, 256, 1, 8)
        b.add_picture(rec, c)
        b.upload()
        b.reconstruct(hb.STAGE_ALL)
        y = b.read_plane(c, 0)
        b.close()
        r, cidx = divmod(k, 4)
        assert np.array_equal(grid_y[r * 256:(r + 1) * 256, cidx * 256:(cidx + 1) * 256], y), k
