#!/usr/bin/env python
"""Measure the roofline denominators MEASURED_PEAKS.json lacks: fp64 DMMA / DFMA peak per MMA shape and
CTA size, isolated trailing-update (SYRK) kernel throughput vs size, and stream-copy bandwidth."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygps_b200 import _lib, build  # noqa: E402

build.build()
e = _lib.Engine(0)
out = {"dmma": {}, "syrk": {}, "copy_gbs": e.bench_copy(1 << 30, 5)}
names = {0: "m8n8k4", 1: "m16n8k4", 2: "m16n8k8", 3: "m16n8k16", 4: "dfma"}
for shape, name in names.items():
    for warps in (2, 4, 8, 16):
        tf, ms = e.bench_dmma(shape, warps, 4000)
        out["dmma"]["%s_w%d" % (name, warps)] = round(tf, 3)
for n in (2048, 4096, 8192, 16256):
    ms, tf = e.bench_syrk(n, 128, 5)
    out["syrk"]["n%d_k128" % n] = {"ms": round(ms, 4), "tflops": round(tf, 3)}
for k in (256, 512):
    ms, tf = e.bench_syrk(8192, k, 3)
    out["syrk"]["n8192_k%d" % k] = {"ms": round(ms, 4), "tflops": round(tf, 3)}
print(json.dumps(out, indent=1))
