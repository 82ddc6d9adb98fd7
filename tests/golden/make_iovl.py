#!/usr/bin/env python
"""'iovl' (image overlay) fixtures under tests/golden/iovl/ and what the UNMODIFIED reference (oracle/_ref,
heif_decode_image) returns for them in every interleaved format -> tests/golden/iovl.json (with the layout of each file,
which the CPU-side restatement in tests/test_overlay.py composes from).
  python tests/golden/make_iovl.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools import heif_writer as W, hevcenc  # noqa: E402
import refheif as R  # noqa: E402

OUT = os.path.join(HERE, "iovl")
FORMATS = {"rgb": R.CHROMA_RGB, "rgba": R.CHROMA_RGBA, "rrggbb_be": R.CHROMA_RRGGBB_BE, "rrggbbaa_be": R.CHROMA_RRGGBBAA_BE,
           "rrggbb_le": R.CHROMA_RRGGBB_LE, "rrggbbaa_le": R.CHROMA_RRGGBBAA_LE}


def enc(w, h, cf, bd, seed, **kw):
    return hevcenc.encode(hevcenc.synth_image(w, h, cf, bd, seed), chroma_format=cf, bit_depth=bd, seed=seed, **kw)


def build():
    """name -> (file bytes, layout): layout = canvas size, background (16 bit), children [(w, h, dx, dy, has_alpha)]"""
    files = {}
    # three 4:4:4 children: opaque, alpha-blended over the first, clipped at the canvas edge; a fourth lies outside the canvas
    b = W.HeifBuilder()
    kids = [(120, 80, 10, 20, False, dict()), (100, 90, 90, 60, True, dict(full_range=0, matrix=1)), (96, 64, 200, 150, False, dict(matrix=9, primaries=9)),
            (64, 64, 300, 10, False, dict())]
    ids = []
    for k, (w, h, dx, dy, alpha, opts) in enumerate(kids):
        iid = b.add_hevc_image(enc(w, h, 3, 8, 80 + k, **opts), w, h, 3, 8, hidden=True)
        if alpha:
            b.add_alpha(iid, enc(w, h, 0, 8, 90 + k), w, h, 0, 8)
        ids.append(iid)
    b.primary = b.add_overlay(ids, 260, 180, [(dx, dy) for _, _, dx, dy, _, _ in kids], background=(0x1234, 0x8000, 0xffff, 0xffff))
    files["iovl_444_three_children"] = (b.serialize(), {"canvas": [260, 180], "background": [0x1234, 0x8000, 0xffff, 0xffff],
                                                        "children": [[w, h, dx, dy, int(a)] for w, h, dx, dy, a, _ in kids]})
    # a 4:4:4 grid as the only child, 32-bit offset fields
    b = W.HeifBuilder()
    tiles = []
    full = hevcenc.synth_image(128, 128, 3, 8, 99)
    for r in range(2):
        for c in range(2):
            planes = [p[r * 64:(r + 1) * 64, c * 64:(c + 1) * 64] for p in full]
            tiles.append(b.add_hevc_image(hevcenc.encode(planes, chroma_format=3, bit_depth=8, seed=99 + r * 2 + c), 64, 64, 3, 8, hidden=True))
    gid = b.add_grid(tiles, 2, 2, 120, 100)
    b.items[[it["id"] for it in b.items].index(gid)]["hidden"] = True
    b.primary = b.add_overlay([gid], 70000, 160, [(40000, 30)], background=(0, 0, 0, 0))
    files["iovl_grid_child_wide_fields"] = (b.serialize(), {"canvas": [70000, 160], "background": [0, 0, 0, 0], "children": [[120, 100, 40000, 30, 0]]})
    return files


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = {}
    for name, (data, layout) in sorted(build().items()):
        open(os.path.join(OUT, name + ".heic"), "wb").write(data)
        m = dict(layout)
        for key, chroma in FORMATS.items():
            try:
                r = R.decode(data, R.COLORSPACE_RGB, chroma)["interleaved"]
                m[key + "_md5"] = hashlib.md5(r[0]).hexdigest()
                m["size"] = [r[1], r[2]]
            except RuntimeError as e:
                m[key + "_error"] = str(e)
        meta[name] = m
        print(name, {k: v for k, v in m.items() if k != "children"})
    json.dump(meta, open(os.path.join(HERE, "iovl.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
