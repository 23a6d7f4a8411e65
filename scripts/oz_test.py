"""int8 tensor-core (Ozaki) trailing update vs the DMMA path vs numpy: accuracy and time."""
import sys, numpy as np
sys.path.insert(0, ".")
from pygps_b200._lib import Engine

eng = Engine()
rng = np.random.default_rng(0)
import os
for n, kw in ([(int(os.environ.get("OZ_N", 15616)), int(os.environ.get("OZ_KW", 384)))] if os.environ.get("OZ_BIG") else [(256, 128), (1024, 384), (4096, 384), (15616, 384)]):
    P = rng.standard_normal((n, kw)) * np.exp(rng.uniform(-6, 6, size=(n, 1)))   # rows of very different scale
    if n <= 4096:
        C = rng.standard_normal((n, n)); C = C + C.T
        ref = C - P @ P.T
        den = np.abs(P) @ np.abs(P).T
        for mode in (0, 1):
            out, ms = eng.dbg_oz_syrk(P, C, mode=mode, reps=3)
            L = np.tril_indices(n)
            err = np.max(np.abs(out[L] - ref[L]) / den[L])
            untouched = np.array_equal(np.triu(out, 1)[::64, ::64], np.triu(C, 1)[::64, ::64]) if mode == 0 else True
            print(f"n={n} kw={kw} mode={mode} max err / (|P||P|') = {err:.3e}  ms={ms:.3f} upper-untouched={untouched}", flush=True)
    else:
        C = np.zeros((n, n), order="F")
        for mode in (0, 1):
            out, ms = eng.dbg_oz_syrk(P, C, mode=mode, reps=5)
            tf = kw * n * n / ms / 1e9
            print(f"n={n} kw={kw} mode={mode} ms={ms:.3f}  fp64-equivalent {tf:.1f} TFLOP/s", flush=True)
            if mode == 0: o0 = out
        print("max |oz - dmma| rel:", np.max(np.abs(np.tril(o0 - out))) / np.max(np.abs(out)))
