#!/usr/bin/env python
"""Builds tools/csc_pipeline_probe.cc against the unmodified reference (oracle/_ref) and records, for every input class x
interleaved output format, the op chain ColorConversionPipeline::construct_pipeline picks -> tests/golden/csc_pipelines.json.
  python tests/golden/make_csc_pipelines.py"""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
LIB = os.path.join(ROOT, "oracle", "_ref")

with tempfile.TemporaryDirectory() as tmp:
    exe = os.path.join(tmp, "probe")
    subprocess.check_call(["g++", "-std=gnu++11", "-w", "-I" + os.path.join(LIB, "gen"), "-I" + os.path.join(LIB, "gen", "libheif"), "-I" + REF,
                           "-I" + os.path.join(REF, "libheif"), "-I" + os.path.join(REF, "libheif", "api"),
                           os.path.join(ROOT, "tools", "csc_pipeline_probe.cc"), "-o", exe, "-L" + LIB, "-lheifref", "-lde265ref", "-Wl,-rpath," + LIB])
    for arg, name in (([], "csc_pipelines.json"), (["bilinear"], "csc_pipelines_bilinear.json")):
        rows = json.loads(subprocess.check_output([exe] + arg))
        # compact form: one string per case
        out = ["%d %d %d %d %d %d %s" % (r["chroma"], r["depth"], r["full"], r["matrix"], r["alpha"], r["out"], r["ops"]) for r in rows]
        json.dump(out, open(os.path.join(HERE, name), "w"), indent=0)
        print(name, len(out), "cases;", len({r["ops"] for r in rows}), "distinct chains")
