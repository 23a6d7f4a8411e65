#!/usr/bin/env python
"""Throughput of S concurrent evaluation streams on ONE GPU (S handles, S threads; independent hyper-parameter vectors,
as the optimizers' random restarts are): python scripts/concurrent_eval.py [N] [evals] [S ...]"""
import math, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
from pygps_b200._dist import replica_hyp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
Ss = [int(v) for v in sys.argv[3:]] or [1, 2, 3]
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8)); y = (np.sin(X.sum(1)) + 0.1 * rng.standard_normal(N))
for S in Ss:
    engs = [_lib.Engine(0) for _ in range(S)]
    for e in engs:
        e.set_data(X)
        for k in range(3):
            e.exact_eval(_lib.COV_RBF, 3, *replica_hyp(k, 0), y, False)
    out = [None] * S

    def work(i):
        r = []
        for k in range(K // S):
            h, sn = replica_hyp(10 + k * S + i, 0)
            r.append(engs[i].exact_eval(_lib.COV_RBF, 3, h, sn, y, False)[0])
        out[i] = r
    ths = [threading.Thread(target=work, args=(i,)) for i in range(S)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    n_done = (K // S) * S
    print("streams %d: %d evals in %.3f s = %.2f evals/s (%.2f ms per eval)" % (S, n_done, dt, n_done / dt, 1e3 * dt / n_done), flush=True)
    del engs
