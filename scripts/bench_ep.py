#!/usr/bin/env python
"""Time GPC + inf.EP + lik.Erf (BASELINE config 5 family): python scripts/bench_ep.py N [D] [der]"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
der = len(sys.argv) > 3 and sys.argv[3] == "der"
rng = np.random.default_rng(0)
X = rng.standard_normal((N, D))
lab = np.sign(X[:, :1] + 0.5 * X[:, 1:2] + 0.3 * rng.standard_normal((N, 1))); lab[lab == 0] = 1
e = _lib.Engine(0)
e.set_data(X)
t0 = time.perf_counter()
out = e.ep_eval(_lib.COV_RBF, 3, [math.log(4.0), 0.0], np.zeros(N), lab, None, None, False, der)
wall = time.perf_counter() - t0
st = e.stats()
print(json.dumps({"config": "GPC EP, cov.RBF, N=%d D=%d fp64, 1 GPU%s" % (N, D, ", der" if der else ""),
                  "nlZ": float(out[0]), "sweeps": out[7], "ms_per_eval": st["total_ms"], "wall_s": wall,
                  "s_per_sweep": st["potrf_ms"] / 1e3 / max(out[7], 1), "launches": st["launches"],
                  "stage_ms": {k: st[k] for k in ("kbuild_ms", "potrf_ms", "solve_ms", "deriv_ms")}}))
