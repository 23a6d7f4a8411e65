#!/usr/bin/env python
"""Time the block-column sharded evaluation (BASELINE config 3 family: RBFard, D=32) under torchrun.

    python -m torch.distributed.run --nproc-per-node G scripts/bench_dist.py N [D] [reps]
Prints one JSON line on rank 0: per-eval device time (max over ranks), Cholesky TFLOP/s (N^3/3), and the
self-consistency residual |(K+sn2 I) alpha - y| / |y| on a sample of rows (the reference cannot run at N=65536)."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
from pygps_b200._dist import DistCtx

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = DistCtx()
eng = _lib.Engine(ctx.local_rank)
ctx.shard_engine(eng)
rng = np.random.default_rng(0)
X = rng.standard_normal((N, D))
y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
hyp = [math.log(3.0)] * D + [0.0]
eng.set_data(X)
times = []
for i in range(reps + 1):
    ctx.barrier()
    nlZ, alpha = eng.exact_eval_dist(_lib.COV_RBFARD, 3, hyp, math.log(0.1), y.reshape(-1))
    st = eng.stats()
    times.append(ctx.max(st["total_ms"]))
    stages = {k: ctx.max(st[k]) for k in ("kbuild_ms", "potrf_ms", "solve_ms")}
best = min(times[1:])
# residual on 256 sampled rows, computed on the host in numpy (self-consistency, not an oracle)
idx = np.random.default_rng(1).choice(N, size=min(256, N), replace=False)
Xs = X / 3.0
d2 = ((Xs[idx, None, :] - Xs[None, :, :]) ** 2).sum(-1) if N <= 8192 else \
    (np.sum(Xs[idx] ** 2, 1)[:, None] + np.sum(Xs ** 2, 1)[None, :] - 2 * Xs[idx] @ Xs.T)
Krows = np.exp(-0.5 * np.maximum(d2, 0))
res = Krows @ alpha + 0.01 * alpha[idx] - y[idx]
if ctx.rank == 0:
    print(json.dumps({"config": "GPR Exact, cov.RBFard, N=%d D=%d fp64, %d GPU(s), block-column sharded" % (N, D, ctx.world),
                      "ms_per_eval": best, "evals_per_s": 1e3 / best, "cholesky_tflops": N ** 3 / 3.0 / (stages["potrf_ms"] * 1e-3) / 1e12,
                      "stage_ms": stages, "nlZ": float(nlZ),
                      "residual_rel": float(np.linalg.norm(res) / np.linalg.norm(y[idx]))}))
ctx.close()
