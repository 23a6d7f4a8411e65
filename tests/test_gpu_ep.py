"""GPC + inf.EP + lik.Erf on the GPU (BASELINE config 5 family) against the reference's frozen outputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import pygps_b200 as pg            # noqa: E402
from oracle import gp_oracle as go  # noqa: E402


from parity_report import check   # noqa: E402

# EP is a fixed-point iteration that the reference stops when nlZ moves by less than 1e-4 between sweeps
# (Core/inf.py:756,761); both implementations run the same number of sweeps from the same start, so they are compared
# at rounding level, not at the iteration's own tolerance.


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def test_kat5_classification_fixture(golden):
    g = golden("classification")
    x, y, xs = g["kat5_x"], g["kat5_y"], g["kat5_xs"]
    m = pg.GPC()
    nlZ, dn, post = m.getPosterior(x, y)
    assert type(nlZ) is np.float64
    check("kat5 nlZ", nlZ, float(g["kat5_nlZ"]), 1e-8)
    assert abs(nlZ - 50.454379530956) < 1e-4
    check("kat5 dcov", dn.cov, g["kat5_dcov"])
    assert dn.lik == [] and dn.mean == []
    assert post.alpha.shape == (120, 1) and post.sW.shape == (120, 1) and post.L.shape == (120, 120)
    check("kat5 alpha", post.alpha, g["kat5_alpha"])
    check("kat5 sW", post.sW, g["kat5_sW"])
    assert np.all(np.tril(post.L, -1) == 0)
    check("kat5 post.L", post.L, g["kat5_L"])
    check("kat5 ttau", m.inffunc.last_ttau, g["kat5_ttau"])
    check("kat5 tnu", m.inffunc.last_tnu, g["kat5_tnu"])
    out = m.predict(xs, np.ones((xs.shape[0], 1)))
    for name, v in zip(("ym", "ys2", "fm", "fs2", "lp"), out):
        check("kat5 " + name, v, g["kat5_" + name])


@pytest.mark.parametrize("N", [200, 512, 1024])
def test_c5_family_synthetic(golden, N):
    """N=1024 is the size of SURVEY 8(c)/(d)'s pin for config 5.  The golden is frozen from the unmodified reference
    with the recipe in oracle/gen_golden.py:c5_big (nlZ 338.71384325684556); SURVEY's printed value (344.43995473)
    came from an input variant its text does not record - the N=512 value of the same table (199.32798736) does
    reproduce with this recipe and is asserted below."""
    g = golden("classification" if N <= 512 else "classification_c5big")
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, 16))
    lab = np.sign(X[:, :1] + 0.5 * X[:, 1:2] + 0.3 * rng.standard_normal((N, 1)))
    lab[lab == 0] = 1
    m = pg.GPC()
    m.setPrior(kernel=pg.cov.RBF(np.log(4.0), 0.0))
    nlZ, dn, post = m.getPosterior(X, lab)
    tag = "c5_%d" % N
    check(tag + " nlZ", nlZ, float(g[tag + "_nlZ"]), 1e-8)
    check(tag + " dcov", dn.cov, g[tag + "_dcov"])
    check(tag + " alpha", post.alpha, g[tag + "_alpha"])
    check(tag + " sW", post.sW, g[tag + "_sW"])
    Xs = np.random.default_rng(1).standard_normal((64, 16))
    out = m.predict(Xs)
    check(tag + " ym", out[0], g[tag + "_ym"])
    check(tag + " fm", out[2], g[tag + "_fm"])
    check(tag + " fs2", out[3], g[tag + "_fs2"])
    assert out[4] is None
    if N == 512:
        assert abs(float(g[tag + "_nlZ"]) - 199.32798736) < 1e-6          # BASELINE.md C5 scaled
    if N == 1024:
        assert abs(float(g[tag + "_nlZ"]) - 338.71384325684556) < 1e-9


def test_warm_start_and_const_mean_match_the_oracle():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((150, 4))
    lab = np.sign(X[:, :1] - 0.3 + 0.2 * rng.standard_normal((150, 1))); lab[lab == 0] = 1
    m = pg.GPC()
    m.setPrior(mean=pg.mean.Const(0.2), kernel=pg.cov.RBFard(log_ell_list=[0.3, 0.1, 0.5, 0.2], log_sigma=0.4))
    nlZ1, dn1, post1 = m.getPosterior(X, lab)
    spec = ("rbfard", [0.3, 0.1, 0.5, 0.2, 0.4])
    rpost, rnlZ, rdn, extra = go.ep_evaluate(("const", 0.2), spec, X, lab, nargout=3)
    check("ep/ard+const nlZ", nlZ1, rnlZ, 1e-8)
    check("ep/ard+const dcov", dn1.cov, rdn["cov"])
    check("ep/ard+const dmean", dn1.mean, rdn["mean"])
    m.covfunc.hyp = [0.35, 0.1, 0.5, 0.2, 0.4]                     # second call warm-starts from the first
    nlZ2, dn2, _ = m.getPosterior(X, lab)
    spec2 = ("rbfard", [0.35, 0.1, 0.5, 0.2, 0.4])
    _, rnlZ2, _, _ = go.ep_evaluate(("const", 0.2), spec2, X, lab, nargout=3, last=(extra["ttau"], extra["tnu"]))
    check("ep/warm start nlZ", nlZ2, rnlZ2)


def test_labels_are_checked_and_shapes_follow_the_reference():
    x = np.random.default_rng(0).standard_normal((20, 2))
    with pytest.raises(Exception):
        pg.GPC().getPosterior(x, np.arange(20.0).reshape(-1, 1))
    y = np.sign(x[:, :1]); y[y == 0] = 1
    post, nlZ, dnlZ = pg.inf.EP().evaluate(pg.mean.Zero(), pg.cov.RBF(), pg.lik.Erf(), x, y, nargout=3)
    assert post.alpha.shape[0] == 20 and post.L.shape == (20, 20) and post.sW.shape == (20, 1)
    assert type(nlZ) is np.float64 and all(type(v) is np.float64 for v in dnlZ.cov)


def test_int8_and_dmma_ep_rebuild_agree(monkeypatch):
    """EP with the two O(n^3) products of the per-sweep rebuild (V = L^-1 sW K as a blocked sweep, Sigma = K - V'V as a
    sliced SYRK) on the int8 tensor cores (default for n >= 2048) against fp64 DMMA (GPK_OZAKI_EP=0)."""
    import math
    from pygps_b200 import _lib
    rng = np.random.default_rng(31)
    n, D = 2300, 4
    X = rng.standard_normal((n, D))
    y = np.sign(X[:, 0] + 0.5 * X[:, 1] + 0.3 * rng.standard_normal(n))
    eng = _lib.Engine(0)
    eng.set_data(X)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GPK_OZAKI_EP", mode)
        res[mode] = eng.ep_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.3], np.zeros(n), y, np.zeros(n), np.zeros(n), False, True)
    a, b = res["1"], res["0"]
    assert abs(a[0] - b[0]) <= 1e-9 * abs(b[0]), (a[0], b[0])
    for u, v in zip(a[1:], b[1:]):
        if isinstance(u, np.ndarray) and u.dtype.kind == "f" and u.size:
            assert np.max(np.abs(u - v)) <= 1e-6 * max(np.max(np.abs(v)), 1e-300)
