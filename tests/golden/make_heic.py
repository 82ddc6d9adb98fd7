#!/usr/bin/env python
"""Generates the HEIC file fixtures under tests/golden/heic/ (tools/heif_writer + tools/hevc_enc) and
records what the UNMODIFIED reference (oracle/_ref/libheifref.so, heif_decode_image) returns for them.

  python tests/golden/make_heic.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools import heif_writer, hevcenc  # noqa: E402
import refheif as R  # noqa: E402

OUT = os.path.join(HERE, "heic")


def md5(b):
    return hashlib.md5(b).hexdigest()


def enc(w, h, cf, bd, seed, **kw):
    return hevcenc.encode(hevcenc.synth_image(w, h, cf, bd, seed), chroma_format=cf, bit_depth=bd, seed=seed, **kw)


def build():
    files = {}
    # grids (BASELINE config C2 in miniature): full-range tiles paste by memcpy, integer CSC
    files["grid_300x200_t128"] = heif_writer.synth_grid_heic(300, 200, tile=128, seed=3)
    files["grid_640x384_t128_wpp"] = heif_writer.synth_grid_heic(640, 384, tile=128, seed=4, wpp=1, log2_ctb=5)
    # tiles without VUI are limited range for the reference: float rescale while pasting (context.cc:2504)
    files["grid_limited_260x130_t64"] = heif_writer.synth_grid_heic(260, 130, tile=64, seed=5, vui=0)
    files["grid_709_limited"] = heif_writer.synth_grid_heic(256, 128, tile=64, seed=6, vui=1, full_range=0, matrix=1)
    files["grid_444"] = heif_writer.synth_grid_heic(200, 120, tile=64, seed=7, chroma_format=3)
    files["grid_422_10"] = heif_writer.synth_grid_heic(200, 120, tile=64, seed=8, chroma_format=2, bit_depth=10)
    # single images
    files["single_420_8_full601"] = heif_writer.single_image(enc(264, 200, 1, 8, 9), 264, 200, 1, 8)
    files["single_420_8_limited709"] = heif_writer.single_image(enc(264, 200, 1, 8, 10, full_range=0, matrix=1), 264, 200, 1, 8)
    files["single_420_8_novui"] = heif_writer.single_image(enc(264, 200, 1, 8, 11, vui=0), 264, 200, 1, 8)
    files["single_420_8_colr_override"] = heif_writer.single_image(enc(264, 200, 1, 8, 12, vui=0), 264, 200, 1, 8, nclx=(1, 13, 1, 1))
    files["single_422_8"] = heif_writer.single_image(enc(200, 120, 2, 8, 13), 200, 120, 2, 8)
    files["single_444_8_bt2020"] = heif_writer.single_image(enc(200, 120, 3, 8, 14, matrix=9, primaries=9), 200, 120, 3, 8)
    files["single_mono_8"] = heif_writer.single_image(enc(200, 120, 0, 8, 15), 200, 120, 0, 8)
    files["single_420_10"] = heif_writer.single_image(enc(200, 120, 1, 10, 16), 200, 120, 1, 10)
    files["single_420_12_limited"] = heif_writer.single_image(enc(200, 120, 1, 12, 17, full_range=0, matrix=9), 200, 120, 1, 12)
    # alpha auxiliary images (BASELINE config C3 in miniature)
    files["alpha_420_8"] = heif_writer.single_image(enc(200, 120, 1, 8, 18), 200, 120, 1, 8, alpha_stream=enc(200, 120, 0, 8, 19))
    files["alpha_422_10"] = heif_writer.single_image(enc(200, 120, 2, 10, 20), 200, 120, 2, 10, alpha_stream=enc(200, 120, 0, 10, 21))
    files["alpha_422_12"] = heif_writer.single_image(enc(200, 120, 2, 12, 22), 200, 120, 2, 12, alpha_stream=enc(200, 120, 0, 12, 23))
    files["alpha_444_8_limited"] = heif_writer.single_image(enc(200, 120, 3, 8, 24, full_range=0), 200, 120, 3, 8,
                                                            alpha_stream=enc(200, 120, 1, 8, 25), alpha_chroma_format=1)
    # geometric transformations: irot / imir on single images, grids, odd sizes, alpha, 10 bit
    W = heif_writer
    files["irot90_420_8_odd"] = W.single_image(enc(263, 199, 1, 8, 30), 263, 199, 1, 8, transforms=(W.irot(1),))
    files["irot180_420_8"] = W.single_image(enc(264, 200, 1, 8, 31), 264, 200, 1, 8, transforms=(W.irot(2),))
    files["irot270_444_8"] = W.single_image(enc(200, 120, 3, 8, 32), 200, 120, 3, 8, transforms=(W.irot(3),))
    files["imir_h_420_8_odd"] = W.single_image(enc(263, 199, 1, 8, 33), 263, 199, 1, 8, transforms=(W.imir(1),))
    files["imir_v_422_8"] = W.single_image(enc(200, 120, 2, 8, 34), 200, 120, 2, 8, transforms=(W.imir(0),))
    files["irot90_imir_420_8"] = W.single_image(enc(264, 200, 1, 8, 35), 264, 200, 1, 8, transforms=(W.irot(1), W.imir(1)))
    files["imir_irot270_420_8"] = W.single_image(enc(264, 200, 1, 8, 36), 264, 200, 1, 8, transforms=(W.imir(0), W.irot(3)))
    files["grid_irot90_300x200"] = W.synth_grid_heic(300, 200, tile=128, seed=37, transforms=(W.irot(1),))
    files["grid_irot270_imir_301x199"] = W.synth_grid_heic(301, 199, tile=128, seed=38, transforms=(W.irot(3), W.imir(1)))
    files["irot90_420_10"] = W.single_image(enc(200, 120, 1, 10, 39), 200, 120, 1, 10, transforms=(W.irot(1),))
    files["alpha_irot90_420_8"] = W.single_image(enc(200, 120, 1, 8, 40), 200, 120, 1, 8, alpha_stream=enc(200, 120, 0, 8, 41),
                                                 transforms=(W.irot(1),))
    # clean aperture (clap): centred and off-centre windows, fractional values, after / before a rotation, 4:2:0 odd windows
    files["clap_centre_420_8"] = W.single_image(enc(264, 200, 1, 8, 50), 264, 200, 1, 8, transforms=(W.clap(200, 1, 120, 1, 0, 1, 0, 1),))
    files["clap_offset_420_8"] = W.single_image(enc(264, 200, 1, 8, 51), 264, 200, 1, 8, transforms=(W.clap(101, 1, 77, 1, -31, 2, 17, 2),))
    files["clap_frac_444_8"] = W.single_image(enc(200, 120, 3, 8, 52), 200, 120, 3, 8, transforms=(W.clap(301, 2, 201, 4, 7, 3, -5, 3),))
    files["irot90_clap_420_8"] = W.single_image(enc(264, 200, 1, 8, 53), 264, 200, 1, 8, transforms=(W.irot(1), W.clap(150, 1, 201, 1, 3, 1, -10, 1)))
    files["clap_irot270_420_8"] = W.single_image(enc(264, 200, 1, 8, 54), 264, 200, 1, 8, transforms=(W.clap(151, 1, 99, 1, 5, 1, 4, 1), W.irot(3)))
    files["grid_clap_imir_300x200"] = W.synth_grid_heic(300, 200, tile=128, seed=55, transforms=(W.clap(255, 1, 131, 1, -11, 1, 9, 1), W.imir(1)))
    files["clap_422_10"] = W.single_image(enc(200, 120, 2, 10, 56), 200, 120, 2, 10, transforms=(W.clap(99, 1, 61, 1, 10, 1, 0, 1),))
    # alpha image of another size / with other transformations than the colour image: nearest-neighbour rescale
    # (context.cc:2064-2071, pixelimage.cc:1156-1253)
    files["alpha_half_420_8"] = W.single_image(enc(200, 120, 1, 8, 60), 200, 120, 1, 8, alpha_stream=enc(100, 60, 0, 8, 61), alpha_size=(100, 60))
    files["alpha_odd_422_10"] = W.single_image(enc(200, 120, 2, 10, 62), 200, 120, 2, 10, alpha_stream=enc(136, 72, 0, 10, 63), alpha_size=(136, 72))
    files["alpha_larger_420_8_irot90"] = W.single_image(enc(200, 120, 1, 8, 64), 200, 120, 1, 8, alpha_stream=enc(264, 200, 1, 8, 65),
                                                        alpha_chroma_format=1, alpha_size=(264, 200), transforms=(W.irot(1),))
    files["alpha_unrotated_420_8"] = W.single_image(enc(200, 120, 1, 8, 66), 200, 120, 1, 8, alpha_stream=enc(200, 120, 0, 8, 67),
                                                    transforms=(W.irot(1),), alpha_transforms=())
    # bit-depth changing conversions and chromaticity-derived matrices: monochrome deeper than
    # 8 bit (Op_mono_to_YCbCr420 + range handling), 4:4:4 12 bit limited, 4:2:2 10 bit full, matrix_coefficients 12 with
    # BT.2020 / BT.709 primaries
    files["single_mono_10_novui"] = heif_writer.single_image(enc(200, 120, 0, 10, 70, vui=0), 200, 120, 0, 10)
    files["single_mono_12_full"] = heif_writer.single_image(enc(200, 120, 0, 12, 71), 200, 120, 0, 12)
    files["single_444_12_limited"] = heif_writer.single_image(enc(200, 120, 3, 12, 72, full_range=0, matrix=1), 200, 120, 3, 12)
    files["single_422_10_full"] = heif_writer.single_image(enc(200, 120, 2, 10, 73), 200, 120, 2, 10)
    files["single_420_8_matrix12_bt2020"] = heif_writer.single_image(enc(200, 120, 1, 8, 74, matrix=12, primaries=9), 200, 120, 1, 8)
    files["single_444_10_matrix13_bt709"] = heif_writer.single_image(enc(200, 120, 3, 10, 75, matrix=13, primaries=1, full_range=0), 200, 120, 3, 10)
    # alpha images that are grids themselves (decode_image_planar of the alpha item, context.cc:2040-2071)
    def alpha_grid(colour, cw, ch, aw, ah, tile, seed):
        b = W.HeifBuilder()
        cols, rows = (aw + tile - 1) // tile, (ah + tile - 1) // tile
        tiles = [enc(tile, tile, 0, 8, seed + k) for k in range(rows * cols)]
        b.primary = colour(b)
        b.add_alpha_grid(b.primary, tiles, rows, cols, tile, tile, aw, ah, 8)
        return b.serialize()
    files["alpha_grid_single_420_8"] = alpha_grid(lambda b: b.add_hevc_image(enc(200, 120, 1, 8, 90), 200, 120, 1, 8), 200, 120, 200, 120, 64, 91)
    files["alpha_grid_half_420_8"] = alpha_grid(lambda b: b.add_hevc_image(enc(200, 120, 1, 8, 100), 200, 120, 1, 8), 200, 120, 100, 60, 64, 101)
    def colour_grid(b):
        tiles = [b.add_hevc_image(enc(128, 128, 1, 8, 110 + k), 128, 128, 1, 8, hidden=True) for k in range(6)]
        return b.add_grid(tiles, 2, 3, 300, 200)
    files["alpha_grid_on_grid_420_8"] = alpha_grid(colour_grid, 300, 200, 300, 200, 128, 120)
    # what a phone writes around the picture: thumbnail ('thmb'), a non-alpha auxiliary image (HDR gain map URN), an Exif
    # item, an ICC profile ('colr' prof) and pixi — none of it may disturb the decode of the primary grid
    def phone_like():
        import struct as _st
        b = W.HeifBuilder()
        tiles = [b.add_hevc_image(enc(128, 128, 1, 8, 130 + k), 128, 128, 1, 8, hidden=True) for k in range(6)]
        icc = W.box(b"colr", b"prof" + bytes(range(64)) * 3)
        pixi = W.fullbox(b"pixi", 0, 0, bytes([3, 8, 8, 8]))
        gid = b.add_grid(tiles, 2, 3, 300, 200, extra_props=(icc, pixi, W.irot(1)))
        thumb = b.add_hevc_image(enc(64, 48, 1, 8, 140), 64, 48, 1, 8, hidden=True)
        b.refs.append((b"thmb", thumb, [gid]))
        auxc = W.fullbox(b"auxC", 0, 0, b"urn:com:apple:photo:2020:aux:hdrgainmap\x00")
        gain = b.add_hevc_image(enc(152, 104, 0, 8, 141), 152, 104, 0, 8, hidden=True, extra_props=(auxc,))
        b.refs.append((b"auxl", gain, [gid]))
        exif = b.add_item(b"Exif", _st.pack(">I", 6) + b"Exif\x00\x00MM\x00*" + bytes(40), [], hidden=True)
        b.refs.append((b"cdsc", exif, [gid]))
        b.primary = gid
        return b.serialize()
    files["phone_like_grid_irot90"] = phone_like()
    files["alpha_prem_420_8"] = heif_writer.single_image(enc(200, 120, 1, 8, 77), 200, 120, 1, 8, alpha_stream=enc(200, 120, 0, 8, 78), premultiplied=True)
    files["single_420_8_gbr"] = heif_writer.single_image(enc(200, 120, 1, 8, 76, matrix=0), 200, 120, 1, 8)
    return files


ALL_FORMATS = {"rgb": R.CHROMA_RGB, "rgba": R.CHROMA_RGBA, "rrggbb_be": R.CHROMA_RRGGBB_BE, "rrggbbaa_be": R.CHROMA_RRGGBBAA_BE,
               "rrggbb_le": R.CHROMA_RRGGBB_LE, "rrggbbaa_le": R.CHROMA_RRGGBBAA_LE}


def reference_all_formats(data, bilinear=False):
    """heif_decode_image(..., heif_colorspace_RGB, chroma) of the unmodified reference for every interleaved chroma;
    bilinear: with the colour conversion options heif-dec -C bilinear sets"""
    out = {}
    for name, chroma in ALL_FORMATS.items():
        try:
            out[name + "_md5"] = md5(R.decode(data, R.COLORSPACE_RGB, chroma, bilinear=bilinear)["interleaved"][0])
        except RuntimeError as e:
            out[name + "_error"] = str(e)
    return out


def reference_outputs(data):
    planes = R.decode(data, R.COLORSPACE_UNDEFINED, R.CHROMA_UNDEFINED)
    bpp = planes["bpp"]
    out = {"bit_depth": bpp, "planes_md5": md5(b"".join(planes[k][0] for k in ("Y", "Cb", "Cr", "A") if k in planes)),
           "width": planes["Y"][1], "height": planes["Y"][2], "has_alpha": "A" in planes}
    targets = {"rgb": R.CHROMA_RGB, "rgba": R.CHROMA_RGBA} if bpp == 8 else {"rrggbb_le": R.CHROMA_RRGGBB_LE, "rrggbbaa_le": R.CHROMA_RRGGBBAA_LE}
    for name, chroma in targets.items():
        try:
            r = R.decode(data, R.COLORSPACE_RGB, chroma)
            out[name + "_md5"] = md5(r["interleaved"][0])
        except RuntimeError as e:
            out[name + "_error"] = str(e)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = {}
    only = set(sys.argv[1:])   # names to (re)generate; default: everything
    if only:
        meta = json.load(open(os.path.join(HERE, "heic.json")))
    for name, data in sorted(build().items()):
        if only and name not in only:
            continue
        open(os.path.join(OUT, name + ".heic"), "wb").write(data)
        meta[name] = reference_outputs(data)
        meta[name]["bytes"] = len(data)
        print(name, meta[name])
    json.dump(meta, open(os.path.join(HERE, "heic.json"), "w"), indent=1, sort_keys=True)
    # every fixture (read back from the committed files) x every interleaved output format, bit-depth changing ones included
    fmts = {}
    for name in sorted(meta):
        fmts[name] = reference_all_formats(open(os.path.join(OUT, name + ".heic"), "rb").read())
    for name in sorted(meta):        # RGBA run through heif_image_rgba_premultiply_alpha
        try:
            data = open(os.path.join(OUT, name + ".heic"), "rb").read()
            fmts[name]["rgba_premultiplied_md5"] = md5(R.decode(data, R.COLORSPACE_RGB, R.CHROMA_RGBA, premultiply=True)["interleaved"][0])
        except RuntimeError as e:
            fmts[name]["rgba_premultiplied_error"] = str(e)
    json.dump(fmts, open(os.path.join(HERE, "heic_formats.json"), "w"), indent=1, sort_keys=True)
    bil = {}
    for name in sorted(meta):
        bil[name] = reference_all_formats(open(os.path.join(OUT, name + ".heic"), "rb").read(), bilinear=True)
    json.dump(bil, open(os.path.join(HERE, "heic_formats_bilinear.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
