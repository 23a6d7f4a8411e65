"""Dense linear-algebra helpers of the hot path, on the GPU.

`jitchol` and `solve_chol` keep the signatures and error contract of
/root/reference/pyGPs/Core/tools.py:31-97 but run in libgpk.so (blocked DMMA Cholesky;
no LAPACK, no LU on a triangular matrix).
"""
import numpy as np

from . import _lib


class _DeviceFactor(np.ndarray):
    """The lower factor returned by jitchol; remembers the engine that still holds it so a
    following solve_chol(L.T, B) can reuse the resident factor instead of re-uploading."""
    _engine = None
    _epoch = -1

    def __array_finalize__(self, obj):
        self._engine = getattr(obj, '_engine', None)
        self._epoch = getattr(obj, '_epoch', -1)


def jitchol(A, maxtries=5):
    """Lower Cholesky factor L (A = L L').  Not positive definite -> np.linalg.LinAlgError,
    as Core/tools.py:62-77 (whose jitter retry is dead code)."""
    A = np.asarray(A, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise Exception('jitchol needs a square matrix')
    eng = _lib.shared_engine()
    R, _ = eng.potrf(A)
    L = np.asfortranarray(R.T).view(_DeviceFactor)        # F-ordered lower factor, like dpotrf's output
    L._engine, L._epoch = eng, eng.epoch
    return L


def solve_chol(L, B):
    """X = A^-1 B given the UPPER factor L of A (A = L'L), Core/tools.py:81-97."""
    try:
        assert(L.shape[0] == L.shape[1] and L.shape[0] == B.shape[0])
    except AssertionError:
        raise Exception('Wrong sizes of matrix arguments in solve_chol.py')
    eng = getattr(L, '_engine', None)
    base = getattr(L, 'base', None)
    if eng is None and base is not None:
        eng = getattr(base, '_engine', None)
        L_epoch = getattr(base, '_epoch', -1)
    else:
        L_epoch = getattr(L, '_epoch', -1)
    if eng is None or eng.epoch != L_epoch:
        # factor not resident: rebuild A's factor on the device from R'R (one SYRK on the host side
        # would defeat the purpose; instead factor A = R'R again, which reproduces R up to rounding)
        R = np.asarray(L, dtype=np.float64)
        eng = _lib.shared_engine()
        eng.potrf(np.dot(R.T, R), want_factor=False)
    return eng.potrs(np.asarray(B, dtype=np.float64))
