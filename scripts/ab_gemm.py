#!/usr/bin/env python
"""A/B the GEMM variants: isolated SYRK rates and one full evaluation each (GPK_GEMM_VARIANT is read at
handle creation, so every variant runs in its own process)."""
import json, math, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, math, os, sys
sys.path.insert(0, %r)
import numpy as np
from pygps_b200 import _lib
e = _lib.Engine(0)
out = {}
for n, k in ((4096, 128), (8192, 128), (16256, 128), (16128, 256), (8192, 512)):
    ms, tf = e.bench_syrk(n, k, 4)
    out["syrk_n%%d_k%%d" %% (n, k)] = round(tf, 2)
N = 16384
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8)); y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
e.set_data(X); e.set_profile(True)
for i in range(3):
    r = e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0) + 0.01 * i, 0.0], math.log(0.1), y.reshape(-1), False)
st = e.stats()
out["eval_ms"] = round(st["total_ms"], 2); out["potrf_ms"] = round(st["potrf_ms"], 2)
out["syrk_insitu_tf"] = round(st["syrk_flops"] / st["syrk_ms"] / 1e9, 2); out["solve_ms"] = round(st["solve_ms"], 2)
print(json.dumps(out))
''' % ROOT
for v in sys.argv[1:] or ["0", "1", "2", "3", "4"]:
    env = dict(os.environ, GPK_GEMM_VARIANT=v)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print("variant", v, r.stdout.strip() or r.stderr[-500:])
