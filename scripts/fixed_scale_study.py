#!/usr/bin/env python
"""CPU study for DESIGN section 8 item 2: slicing the Cholesky panels with FIXED row scales 2^ceil(log2 sqrt(A_ii))
(known before the factorisation starts, since |L_ij| <= sqrt(A_ii)) instead of the per-block row maxima.
Runs the numpy model of the device algorithm (tests/blockref.py) on a C2-like matrix and prints the errors of the factor,
of log det and of the solve against LAPACK for: fp64 updates, per-block scales (what the device does), fixed scales.
    python scripts/fixed_scale_study.py [N] [S] [sn]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import blockref as br

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
S = int(sys.argv[2]) if len(sys.argv) > 2 else 7
SN = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8))
y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(N)
d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
A0 = np.exp(-0.5 * d2 / 4.0) / SN ** 2 + np.eye(N)       # K / sn2 + I
Lref = np.linalg.cholesky(A0)
xref = np.linalg.solve(A0, y)
fixed_e = None


def split_fixed(P, S_=7, RB=8, rows=None):
    """oz_split with the exponent taken from the ORIGINAL diagonal of the matrix (global rows `rows`)."""
    e = fixed_e[rows]
    x = np.ldexp(P, -e[:, None])
    assert np.abs(x).max() <= 0.5
    q = np.rint(x * 2.0 ** (S_ * RB)).astype(np.int64)
    d = np.empty((S_,) + P.shape, dtype=np.int8)
    for t in range(S_ - 1, -1, -1):
        lo = q & ((1 << RB) - 1)
        dd = np.where(lo >= (1 << (RB - 1)), lo - (1 << RB), lo)
        d[t] = dd
        q = (q - dd) >> RB
    assert (q == 0).all()
    return d, np.ldexp(1.0, e - RB)


def run(mode):
    A = br.pad_spd(A0)
    if mode == "fp64":
        br.potrf_device(A, W=3, W1=6, w1_minrem=2, oz=False)
    else:
        orig = br.oz_split
        if mode == "fixed":
            state = {"row0": 0}
            real_syrk = br.oz_syrk

            def syrk(C, P, S_=S, RB=8):
                # the model calls oz_syrk(full, pan) with pan = A[rows0*NB:, k0:k1]; recover rows0 from the height
                rows = np.arange(A.shape[0] - P.shape[0], A.shape[0])
                br.oz_split = lambda PP, SS=S_, RR=RB: split_fixed(PP, SS, RR, rows)
                try:
                    real_syrk(C, P, S_, RB)
                finally:
                    br.oz_split = orig
            br.oz_syrk = syrk
            try:
                br.potrf_device(A, W=3, W1=6, w1_minrem=2, oz=True)
            finally:
                br.oz_syrk = real_syrk
        else:
            br.oz_split = (lambda PP, SS=S, RR=8: orig(PP, S, RR))
            try:
                br.potrf_device(A, W=3, W1=6, w1_minrem=2, oz=True)
            finally:
                br.oz_split = orig
    L = np.tril(A[:N, :N])
    x = np.linalg.solve(L.T, np.linalg.solve(L, y))
    return (np.max(np.abs(L - Lref)) / np.max(np.abs(Lref)), abs(np.log(np.diag(L)).sum() - np.log(np.diag(Lref)).sum()),
            np.max(np.abs(x - xref)) / np.max(np.abs(xref)), np.linalg.norm(A0 @ x - y) / np.linalg.norm(y))


np_ = br.pad_spd(A0).shape[0]
diag = np.ones(np_); diag[:N] = np.diag(A0)
_, ex = np.frexp(np.sqrt(diag))
fixed_e = (ex + 1).astype(np.int64)                      # |L_ij| <= sqrt(A_ii) < 2^ex  ->  |x| <= 0.5
print("N=%d, S=%d digits, cond(A) = %.2e" % (N, S, np.linalg.cond(A0)))
print("%-28s %12s %12s %12s %12s" % ("updates", "max|dL|/|L|", "|d logdet|", "max|dx|/|x|", "residual"))
for mode in ("fp64", "per-block scales (device)", "fixed"):
    r = run("fixed" if mode == "fixed" else ("fp64" if mode == "fp64" else "block"))
    print("%-28s %12.3e %12.3e %12.3e %12.3e" % (((mode if mode != "fixed" else "fixed scales sqrt(A_ii)"),) + tuple(r)))
