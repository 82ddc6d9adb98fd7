#!/usr/bin/env python
"""Generates the synthetic bitstream fixtures under tests/golden/generated/ with the test-content
generator (tools/hevc_enc) and records the UNMODIFIED reference decoder's output MD5 for each
(oracle/_ref/dec265, default SIMD build; "-0" = its scalar build where the two disagree).

  python tests/golden/make_generated.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tools import hevcenc  # noqa: E402

OUT = os.path.join(HERE, "generated")
DEC265 = os.path.join(ROOT, "oracle", "_ref", "dec265")

# name -> (width, height, encoder options)
CASES = {
    "base_420_8":        (264, 200, {}),
    "mono_8":            (200, 120, dict(chroma_format=0)),
    "c422_8":            (200, 120, dict(chroma_format=2)),
    "c444_8":            (200, 120, dict(chroma_format=3)),
    "c420_10":           (200, 120, dict(bit_depth=10)),
    "c422_10":           (264, 200, dict(bit_depth=10, chroma_format=2)),
    "c422_12":           (264, 200, dict(bit_depth=12, chroma_format=2)),
    "c444_10":           (200, 120, dict(bit_depth=10, chroma_format=3)),
    "c420_12":           (200, 120, dict(bit_depth=12)),
    "mono_12":           (200, 120, dict(bit_depth=12, chroma_format=0)),
    "ctb16_nosao":       (264, 200, dict(log2_ctb=4, sao=0)),
    "ctb16_sao":         (264, 200, dict(log2_ctb=4)),   # reference AVX2 SAO bug: golden from the scalar build
    "ctb32_tb16":        (264, 200, dict(log2_ctb=5, log2_max_tb=4)),
    "qp10":              (264, 200, dict(qp=10)),
    "qp40":              (264, 200, dict(qp=40)),
    "qp51":              (200, 120, dict(qp=51)),
    "cuqpdelta_d0":      (264, 200, dict(cu_qp_delta=1)),
    "cuqpdelta_d2":      (264, 200, dict(cu_qp_delta=3)),
    "tskip":             (264, 200, dict(transform_skip=1)),
    "wpp":               (264, 200, dict(wpp=1)),
    "tiles_2x2":         (264, 200, dict(tile_cols=2, tile_rows=2)),
    "tiles_3x2_nolf":    (264, 200, dict(tile_cols=3, tile_rows=2, loop_filter_across_tiles=0, log2_ctb=5)),
    "wpp_slices":        (264, 200, dict(wpp=1, slice_ctbs=10)),
    "slices":            (264, 200, dict(log2_ctb=5, slice_ctbs=13)),
    "slices_dep":        (264, 200, dict(log2_ctb=5, slice_ctbs=13, dependent_slices=1)),
    "slices_nolf":       (264, 200, dict(log2_ctb=5, slice_ctbs=13, loop_filter_across_slices=0)),
    "slices_nolf_444":   (264, 200, dict(log2_ctb=5, slice_ctbs=7, loop_filter_across_slices=0, chroma_format=3)),
    "slices_nolf_422_t": (264, 200, dict(log2_ctb=5, slice_ctbs=7, loop_filter_across_slices=0, chroma_format=2, tile_cols=2)),
    "wpp_dep":           (264, 200, dict(wpp=1, log2_ctb=5, slice_ctbs=18, dependent_slices=1)),
    "scaling_default":   (264, 200, dict(scaling_list=1)),
    "scaling_custom":    (264, 200, dict(scaling_list=2)),
    "scaling_custom_10": (200, 120, dict(scaling_list=2, bit_depth=10, chroma_format=2)),
    "pcm":               (264, 200, dict(pcm=1)),
    "bypass":            (264, 200, dict(transquant_bypass=1)),
    "bypass_10":         (264, 200, dict(transquant_bypass=1, bit_depth=10)),
    "pcm_bypass_422":    (264, 200, dict(pcm=1, transquant_bypass=1, chroma_format=2)),
    "nodeblock":         (264, 200, dict(deblock_disable=1)),
    "dbk_offsets":       (264, 200, dict(beta_offset_div2=3, tc_offset_div2=-2)),
    "chroma_qp_offsets": (264, 200, dict(cb_qp_offset=5, cr_qp_offset=-4)),
    "plain":             (264, 200, dict(strong_intra=0, sao=0, sign_hiding=0)),
    "novui":             (264, 200, dict(vui=0)),
    "odd_size":          (250, 130, dict(seed=3)),
    "odd_size_444":      (251, 131, dict(seed=3, chroma_format=3)),
    "tiny":              (8, 8, dict(seed=4)),
    "one_ctb_wide":      (40, 300, dict(wpp=1, seed=5)),
    "deep_rqt_520":      (520, 520, dict(max_th_depth=3, seed=7)),
    "tile_512":          (512, 512, dict(wpp=1, seed=11)),
}
SCALAR_GOLDEN = {"ctb16_sao"}


def ref_md5(annexb, scalar=False):
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "s.265"), os.path.join(d, "o.yuv")
        open(src, "wb").write(annexb)
        subprocess.check_call([DEC265, "-q", "-t", "0"] + (["-0"] if scalar else []) + ["-o", dst, src],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return hashlib.md5(open(dst, "rb").read()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = {}
    for name, (w, h, kw) in sorted(CASES.items()):
        cf, bd = kw.get("chroma_format", 1), kw.get("bit_depth", 8)
        planes = hevcenc.synth_image(w, h, cf, bd, seed=kw.get("seed", 1))
        stream = hevcenc.encode(planes, **kw)
        open(os.path.join(OUT, name + ".hevc"), "wb").write(stream)
        annexb = hevcenc.to_annexb(stream)
        entry = {"width": w, "height": h, "options": kw, "bytes": len(stream),
                 "yuv_md5": ref_md5(annexb, scalar=name in SCALAR_GOLDEN)}
        if name in SCALAR_GOLDEN:
            entry["reference_simd_md5"] = ref_md5(annexb)
            entry["note"] = ("golden from the reference's scalar build (dec265 -0): its AVX2 SAO kernels overshoot "
                             "8-sample-wide chroma CTBs, so the SIMD build's output depends on the host CPU")
        meta[name] = entry
        print(name, entry["bytes"], entry["yuv_md5"])
    json.dump(meta, open(os.path.join(HERE, "generated.json"), "w"), indent=1, sort_keys=True)
    print("total fixture bytes", sum(e["bytes"] for e in meta.values()))


if __name__ == "__main__":
    main()
