"""Dense linear-algebra helpers of the hot path, on the GPU.

`jitchol` and `solve_chol` keep the signatures and error contract of
/root/reference/pyGPs/Core/tools.py:31-97 but run in libgpk.so (blocked Cholesky with the
trailing updates on the tensor cores; no LAPACK, no LU on a triangular matrix).
"""
import weakref

import numpy as np

from . import _lib

# The factor jitchol returned last: weakref to the array + the memory signature of its transpose.  solve_chol
# reuses the factor still resident on the GPU ONLY for a full-extent transpose view of exactly that (read-only,
# still alive) array; a slice, a copy or anything else is a different matrix and goes through the upload path.
_RESIDENT = {"L": None, "sig": None, "engine": None, "epoch": -1}


class _DeviceFactor(np.ndarray):
    """ndarray subclass only so that the array returned by jitchol can be weakly referenced."""
    pass


def _mem_sig(a):
    return (a.__array_interface__["data"][0], a.shape, a.strides)


def jitchol(A, maxtries=5):
    """Lower Cholesky factor L (A = L L').  Not positive definite -> np.linalg.LinAlgError,
    as Core/tools.py:62-77 (whose jitter retry is dead code)."""
    A = np.asarray(A, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise Exception('jitchol needs a square matrix')
    eng = _lib.shared_engine()
    R, _ = eng.potrf(A)
    L = np.asfortranarray(R.T).view(_DeviceFactor)        # F-ordered lower factor, like dpotrf's output
    # read-only: an in-place edit would silently diverge from the resident factor (copy it to modify it)
    L.flags.writeable = False
    _RESIDENT.update(L=weakref.ref(L), sig=_mem_sig(L.T), engine=eng, epoch=eng.epoch)
    return L


def _resident_engine(R):
    """The engine that still holds exactly this upper factor, or None."""
    r = _RESIDENT
    if r["engine"] is None or r["engine"].epoch != r["epoch"] or r["L"] is None or r["L"]() is None:
        return None
    if R.flags.writeable or _mem_sig(R) != r["sig"]:
        return None
    return r["engine"]


def solve_chol(L, B):
    """X = A^-1 B given the UPPER factor L of A (A = L'L), Core/tools.py:81-97."""
    try:
        assert(L.shape[0] == L.shape[1] and L.shape[0] == B.shape[0])
    except AssertionError:
        raise Exception('Wrong sizes of matrix arguments in solve_chol.py')
    B = np.asarray(B, dtype=np.float64)
    eng = _resident_engine(L) if isinstance(L, np.ndarray) else None
    if eng is None:
        # factor not resident: upload R itself as the factor (gpk_set_factor) - no R'R product, no re-factorisation
        eng = _lib.shared_engine()
        eng.set_factor(np.asarray(L, dtype=np.float64))
    return eng.potrs(B)
