"""Sweep the blocking parameters of the three-level Cholesky: python scripts/sweep_potrf.py N 'W2,W1,MINREM' ..."""
import math, os, sys, time
sys.path.insert(0, ".")
import numpy as np
from pygps_b200 import _lib

N = int(sys.argv[1])
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8))
y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
e = _lib.Engine(0)
e.set_data(X)
e.set_profile(1)
for cfg in sys.argv[2:]:
    w2, w1, mr = cfg.split(",")
    os.environ["GPK_POTRF_W"] = w2; os.environ["GPK_POTRF_W1"] = w1; os.environ["GPK_POTRF_W1_MINREM"] = mr
    ts = []
    for k in range(4):
        out = e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), y.reshape(-1), False)
        ts.append(e.stats()["total_ms"])
    st = e.stats()
    print(f"W2={w2} W1={w1} minrem={mr}: total {min(ts):.2f} ms potrf {st['potrf_ms']:.2f} syrk {st['syrk_ms']:.2f} nlZ {out[0]:.9f}", flush=True)
