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
for (N, K) in [(64, 32), (64, 128), (64, 256), (128, 64)]:
    for mode in (1, 2):
        A = rng.integers(-128, 128, size=(128, K)).astype(np.int8)
        B = rng.integers(-128, 128, size=(N, K)).astype(np.int8)
        C = eng.dbg_i8_tile(A, B, a_tmem=mode)
        ref = A.astype(np.int64) @ B.astype(np.int64).T
        bad = (C != ref)
        print(f"A-in-TMEM mode={mode} N={N} K={K} mismatches={int(bad.sum())} of {C.size}", flush=True)
        if bad.any() and K == 32:
            # layout forensics: which A row/k does each output row actually see?
            for r in (0, 1, 8, 17, 33, 64, 100):
                cand = [(rr) for rr in range(128) if np.array_equal(A[rr].astype(np.int64) @ B.astype(np.int64).T, C[r])]
                print("  out row", r, "matches A row", cand)
