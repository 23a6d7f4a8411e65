#!/usr/bin/env python
"""Run a few exact_eval calls (for ncu launch lists): python scripts/one_eval.py N [reps] [der]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from pygps_b200 import _lib, build  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
der = len(sys.argv) > 3 and sys.argv[3] == "der"
build.build()
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8))
y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
e = _lib.Engine(0)
e.set_data(X)
for k in range(reps):
    out = e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0) + 0.01 * k, 0.0], math.log(0.1), y.reshape(-1), der)
    print("nlZ", out[0], e.stats())
