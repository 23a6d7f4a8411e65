"""CPU checks of the device orchestration logic through its numpy model (tests/blockref.py)."""
import numpy as np
import scipy.linalg as sla

import blockref as br


def _spd(n, seed=0, cond_shift=1.0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, 4))
    d = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    return np.exp(-0.5 * d / 4.0) / 0.01 + cond_shift * np.eye(n)


def test_diag_block_matches_lapack():
    A = _spd(128, 1)
    L, Li, ld, info = br.diag_block(np.asfortranarray(A))
    Lref = np.linalg.cholesky(A)
    assert info == 0
    np.testing.assert_allclose(L, Lref, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(Li @ Lref, np.eye(128), atol=1e-9)
    assert np.all(np.triu(Li, 1) == 0) and np.all(np.triu(L, 1) == 0)
    np.testing.assert_allclose(ld, np.log(np.diag(Lref)).sum(), rtol=1e-13)


def test_diag_block_reports_first_bad_pivot():
    A = _spd(128, 2)
    A[40, 40] = -5.0
    _, _, _, info = br.diag_block(np.asfortranarray(A))
    assert info == 41


import pytest


@pytest.mark.parametrize("n,W", [(300, 1), (300, 2), (700, 2), (700, 3), (520, 4)])
def test_blocked_potrf_with_lookahead_split_and_fused_forward_solve(n, W):
    A = _spd(n, 3)
    rng = np.random.default_rng(0)
    y = rng.standard_normal(n)
    P = br.pad_spd(A)
    b = np.zeros(P.shape[0]); b[:n] = y
    Dinv, parts, info, z = br.potrf_device(P, b, W=W)
    Lref = np.linalg.cholesky(A)
    assert info == 0
    np.testing.assert_allclose(np.tril(P)[:n, :n], Lref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(parts.sum(), np.log(np.diag(Lref)).sum(), rtol=1e-12)
    np.testing.assert_allclose(z[:n], sla.solve_triangular(Lref, y, lower=True), rtol=1e-9, atol=1e-9)
    # backward steps
    T = P.shape[0] // br.NB
    x = np.zeros_like(z)
    for k in range(T - 1, -1, -1):
        br.trsv_bwd(P, Dinv, z, x, k, T)
    np.testing.assert_allclose(x[:n], np.linalg.solve(A, y), rtol=1e-8, atol=1e-8)
    assert np.all(x[n:] == 0)


def test_transposed_forward_sweep_is_predict_solve():
    n, ns = 260, 130
    A = _spd(n, 4)
    P = br.pad_spd(A)
    Dinv, _, _, _ = br.potrf_device(P)
    rng = np.random.default_rng(1)
    Ks = rng.standard_normal((n, ns))
    nsp = 256
    Pt = np.asfortranarray(np.zeros((nsp, P.shape[0])))
    Pt[:ns, :n] = Ks.T
    br.sweep_forward(Pt, P, Dinv, P.shape[0] // br.NB)
    V = sla.solve_triangular(np.linalg.cholesky(A), Ks, lower=True)
    np.testing.assert_allclose(Pt[:ns, :n], V.T, rtol=1e-9, atol=1e-9)


def test_inverse_through_upper_factor():
    n = 384
    A = _spd(n, 5)
    P = br.pad_spd(A)
    Dinv, _, _, _ = br.potrf_device(P)
    U = br.inverse_factor_T(P, Dinv)
    Lref = np.linalg.cholesky(A)
    np.testing.assert_allclose(np.triu(U)[:n, :n], np.linalg.inv(Lref).T, rtol=1e-8, atol=1e-9)
    assert np.all(np.tril(U, -1) == 0)
    W = br.inverse_lower(U)
    Ainv = np.linalg.inv(A)
    np.testing.assert_allclose(np.tril(W)[:n, :n], np.tril(Ainv), rtol=1e-7, atol=1e-9)
