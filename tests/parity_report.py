"""Achieved-error ledger of the GPU parity tests (test infrastructure).

`check(name, got, ref, tol)` asserts max|got-ref| / max|ref| < tol AND records the achieved error, so the terminal
summary (tests/conftest.py) prints every achieved error beside its tolerance - a tolerance looser than the 1e-6 of
BASELINE.json's north_star has to carry a `why` (printed too), normally the condition number it was derived from."""
import json
import os

import numpy as np

LEDGER = []


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def check(name, got, ref, tol=1e-6, why=None):
    err = rel(got, ref)
    if tol > 1e-6 and not why:
        raise AssertionError("%s: tolerance %g is looser than 1e-6 and carries no justification" % (name, tol))
    LEDGER.append({"name": name, "err": err, "tol": tol, "why": why})
    assert err < tol, "%s: achieved %.3e, tolerance %.3e%s" % (name, err, tol, (" (%s)" % why) if why else "")
    return err


def cond_tol(cond, factor=64.0, floor=1e-6, cap=1e-4):
    """Tolerance for a quantity obtained through solves with a matrix of condition number `cond`: two correct fp64
    implementations (the reference's LAPACK path and ours) may differ by O(cond * eps); never looser than `cap`."""
    return float(min(cap, max(floor, factor * cond * np.finfo(float).eps)))


def dump(path):
    if LEDGER:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(LEDGER, f, indent=1)
