#!/usr/bin/env python
"""Time GPR_FITC evaluations (BASELINE config 4 family, fp64) - single GPU or data-sharded under torchrun.

    python scripts/bench_fitc.py N M [D] [reps] [der]
    python -m torch.distributed.run --nproc-per-node G scripts/bench_fitc.py N M ..."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
from pygps_b200._dist import DistCtx
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
D = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
der = len(sys.argv) > 5 and sys.argv[5] == "der"
ctx = DistCtx()
eng = _lib.Engine(ctx.local_rank)
ctx.shard_engine(eng)
rng = np.random.default_rng(0)
X = rng.standard_normal((N, D))
y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
U = rng.standard_normal((M, D))
lo, hi = (N * ctx.rank) // ctx.world, (N * (ctx.rank + 1)) // ctx.world
eng.set_data(X[lo:hi])
ts = []
for i in range(reps + 1):
    ctx.barrier()
    out = eng.fitc_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), U, y[lo:hi].reshape(-1), der)
    st = eng.stats()
    ts.append(ctx.max(st["total_ms"]))
if ctx.rank == 0:
    best = min(ts[1:])
    flops = 2.0 * M * M * N * (1.0 if not der else 1.0) + 2.0 * M ** 3 / 3
    print(json.dumps({"config": "GPR_FITC, cov.RBF, N=%d M=%d D=%d fp64, %d GPU(s)%s" % (N, M, D, ctx.world, ", der" if der else ""),
                      "ms_per_eval": best, "evals_per_s": 1e3 / best, "nlZ": float(out[0]),
                      "tflops_nlz_path": flops / (best * 1e-3) / 1e12 if not der else None,
                      "stage_ms": {k: st[k] for k in ("kbuild_ms", "potrf_ms", "solve_ms", "deriv_ms")}}))
ctx.close()
