import sys, numpy as np
sys.path.insert(0, ".")
from pygps_b200._lib import Engine
eng = Engine()
rng = np.random.default_rng(1)
for (N, K) in [(64, 32), (64, 128), (128, 64), (256, 256)]:
    for (da, db) in [(np.int8, np.int8), (np.uint8, np.int8), (np.int8, np.uint8), (np.uint8, np.uint8)]:
        A = rng.integers(0 if da == np.uint8 else -128, 256 if da == np.uint8 else 128, size=(128, K)).astype(da)
        B = rng.integers(0 if db == np.uint8 else -128, 256 if db == np.uint8 else 128, size=(N, K)).astype(db)
        C = eng.dbg_i8_tile(A, B)
        ref = A.astype(np.int64) @ B.astype(np.int64).T
        print(f"N={N} K={K} A={da.__name__} B={db.__name__} mismatches={int((C != ref).sum())} of {C.size}", flush=True)
