"""Edge cases of the GPU path: tiny and ragged sizes, wide inputs, many test points, repeated use of one handle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import pygps_b200 as pg            # noqa: E402
from oracle import gp_oracle as go  # noqa: E402


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("n,D", [(1, 1), (2, 3), (127, 2), (128, 2), (129, 2), (257, 40), (385, 1)])
def test_ragged_sizes_match_the_oracle(n, D):
    rng = np.random.default_rng(n * 31 + D)
    X = rng.standard_normal((n, D)); y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((n, 1))
    Xs = rng.standard_normal((5, D))
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.RBF(0.2, 0.1))
    nlZ, dn, post = m.getPosterior(X, y)
    rpost, rnlZ, rdn = go.exact_evaluate(("zero",), ("rbf", [0.2, 0.1]), np.log(0.1), X, y, 3)
    assert abs(nlZ - rnlZ) < 1e-9 * max(1.0, abs(rnlZ))
    assert rel(post.alpha, rpost["alpha"]) < 1e-7 and rel(post.L, rpost["L"]) < 1e-9
    assert rel(dn.cov, rdn["cov"]) < 1e-6 and rel(dn.lik, rdn["lik"]) < 1e-6
    out = m.predict(Xs)
    ro = go.predict(("zero",), ("rbf", [0.2, 0.1]), np.log(0.1), X, rpost, Xs)
    assert rel(out[0], ro[0]) < 1e-7 and rel(out[1], ro[1]) < 1e-7


def test_many_test_points_are_chunked():
    rng = np.random.default_rng(2)
    X = rng.standard_normal((300, 3)); y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((300, 1))
    Xs = rng.standard_normal((20000, 3))                      # > the 8192-point device chunk
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.Matern(0.1, 5, 0.2))
    m.getPosterior(X, y, der=False)
    ym, ys2, fm, fs2, lp = m.predict(Xs, np.zeros((20000, 1)))
    rpost, _ = go.exact_evaluate(("zero",), ("matern", [0.1, 0.2], 5), np.log(0.1), X, y, 2)
    ro = go.predict(("zero",), ("matern", [0.1, 0.2], 5), np.log(0.1), X, rpost, Xs, np.zeros((20000, 1)))
    assert rel(ym, ro[0]) < 1e-7 and rel(ys2, ro[1]) < 1e-7 and rel(lp, ro[4]) < 1e-6
    assert ym.shape == (20000, 1) and np.all(fs2 >= 0)


def test_one_model_many_sizes_and_hyperparameters():
    """The handle re-allocates when the problem size changes and never leaks a stale factor."""
    rng = np.random.default_rng(4)
    m = pg.GPR()
    for n in (50, 400, 130, 400):
        X = rng.standard_normal((n, 2)); y = np.cos(X[:, :1]) + 0.05 * rng.standard_normal((n, 1))
        for ell in (0.0, 0.5):
            m.covfunc.hyp = [ell, 0.0]
            nlZ, post = m.getPosterior(X, y, der=False)
            _, rnlZ = go.exact_evaluate(("zero",), ("rbf", [ell, 0.0]), np.log(0.1), X, y, 2)
            assert abs(nlZ - rnlZ) < 1e-9 * max(1.0, abs(rnlZ))
            assert post.L.shape == (n, n)


def test_1d_inputs_and_nan_targets_propagate():
    x = np.linspace(-2, 2, 40); y = np.sin(x)
    m = pg.GPR()
    nlZ, post = m.getPosterior(x, y, der=False)               # 1-d arrays become columns (Core/gp.py:315-323)
    assert post.alpha.shape == (40, 1) and np.isfinite(nlZ)
    ybad = y.copy(); ybad[3] = np.nan
    nlZ2, post2 = m.getPosterior(x, ybad, der=False)          # NaN in y: K is fine, nlZ/alpha become NaN (no abort)
    assert np.isnan(nlZ2) and np.isnan(post2.alpha).any()
