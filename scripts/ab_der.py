#!/usr/bin/env python
"""Derivative evaluation (der=True): int8 inverse / U U' (GPK_OZAKI_DER=1) against the DMMA path (=0)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8)); y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
e = _lib.Engine(0)
e.set_data(X)
res = {}
for mode in ("0", "1", "0", "1"):
    os.environ["GPK_OZAKI_DER"] = mode
    best = 1e9
    for k in range(3):
        t = time.perf_counter()
        out = e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), y.reshape(-1), True)
        best = min(best, time.perf_counter() - t)
    st = e.stats()
    res[mode] = out
    print("N=%d GPK_OZAKI_DER=%s: %.2f ms (deriv stage %.2f ms) nlZ=%.10f dcov=%s dlik=%s" % (
        N, mode, best * 1e3, st["deriv_ms"], out[0], np.array(out[2]), np.array(out[3])), flush=True)
a, b = res["0"], res["1"]
print("rel diff dcov", np.max(np.abs(np.array(a[2]) - np.array(b[2])) / np.abs(np.array(a[2]))),
      "dlik", np.max(np.abs(np.array(a[3]) - np.array(b[3])) / np.abs(np.array(a[3]))))
