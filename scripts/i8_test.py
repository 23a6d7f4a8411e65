import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
e = _lib.Engine(0)
rng = np.random.default_rng(0)
for N, K in ((64, 32), (64, 128), (128, 64), (256, 256)):
    A = rng.integers(-64, 65, size=(128, K), dtype=np.int8)
    B = rng.integers(-64, 65, size=(N, K), dtype=np.int8)
    C = e.dbg_i8_tile(A, B)
    ref = A.astype(np.int32) @ B.astype(np.int32).T
    bad = int((C != ref).sum())
    print("N=%d K=%d mismatches=%d of %d ; C[0,:4]=%s ref=%s" % (N, K, bad, C.size, C[0, :4], ref[0, :4]))
    if bad:
        rows = np.where((C != ref).any(1))[0]; cols = np.where((C != ref).any(0))[0]
        print("  bad rows", rows[:10], "...", len(rows), " bad cols", cols[:10], "...", len(cols))
