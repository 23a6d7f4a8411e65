#!/usr/bin/env python
"""A/B the outer block width W of the two-level Cholesky (GPK_POTRF_W is read on every factorisation)."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
e = _lib.Engine(0)
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8)); y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
e.set_data(X); e.set_profile(True)
for W in (1, 2, 3, 4, 8):
    os.environ["GPK_POTRF_W"] = str(W)
    for i in range(3):
        r = e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), y.reshape(-1), False)
    st = e.stats()
    print("W=%d nlZ=%.10f eval=%.2f ms potrf=%.2f syrk_ms=%.2f insitu=%.2f TF/s solve=%.2f launches=%d" % (
        W, r[0], st["total_ms"], st["potrf_ms"], st["syrk_ms"], st["syrk_flops"] / st["syrk_ms"] / 1e9,
        st["solve_ms"], st["launches"]))
