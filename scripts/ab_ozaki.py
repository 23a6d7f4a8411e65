"""A/B: exact_eval with the trailing update on DMMA (GPK_OZAKI=0) vs the int8 tensor cores (GPK_OZAKI=1)."""
import math, os, sys, time
sys.path.insert(0, ".")
import numpy as np
from pygps_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8))
y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
e = _lib.Engine(0)
e.set_data(X)
e.set_profile(1)
res = {}
for oz in ("0", "1", "0", "1"):
    os.environ["GPK_OZAKI"] = oz
    ts = []
    for k in range(4):
        t = time.perf_counter()
        out = e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), y.reshape(-1), False)
        ts.append(time.perf_counter() - t)
    st = e.stats()
    res[oz] = (out[0], np.array(out[1]).copy())
    print(f"GPK_OZAKI={oz} slices={os.environ.get('GPK_OZAKI_SLICES','8')}: nlZ={out[0]:.12f} wall min {min(ts)*1e3:.2f} ms  stats={st}", flush=True)
a0, a1 = res["0"][1], res["1"][1]
print("nlZ rel diff", abs(res["0"][0] - res["1"][0]) / abs(res["0"][0]), " alpha rel diff (max/|max|)", np.max(np.abs(a0 - a1)) / np.max(np.abs(a0)))
