"""Parity of the drop-in classes against the reference's frozen outputs (tests/golden) and the
CPU oracle, through the plugin API -> C ABI -> CUDA.  Tolerance: north_star's 1e-6 relative
(fp64); most checks are far tighter and say so."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import pygps_b200 as pg            # noqa: E402
from oracle import gp_oracle as go  # noqa: E402

TOL = 1e-6


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _check_model(m, g, tag, x, y, xs, ys=None, der=True, tol=1e-8):
    nlZ, dn, post = m.getPosterior(x, y)
    assert type(nlZ) is np.float64
    assert abs(nlZ - g[tag + "_nlZ"]) <= tol * abs(g[tag + "_nlZ"]), (tag, nlZ, g[tag + "_nlZ"])
    assert post.alpha.shape == (x.shape[0], 1) and post.sW.shape == (x.shape[0], 1)
    assert rel(post.alpha, g[tag + "_alpha"]) < TOL, (tag, "alpha", rel(post.alpha, g[tag + "_alpha"]))
    assert abs(post.sW[0, 0] - g[tag + "_sW0"]) < 1e-12
    if der:
        for got, key in ((dn.cov, "_dcov"), (dn.lik, "_dlik"), (dn.mean, "_dmean")):
            ref = g[tag + key]
            assert len(got) == len(ref)
            if len(ref):
                assert all(type(v) is np.float64 for v in got)
                assert rel(got, ref) < TOL, (tag, key, got, ref)
    if tag + "_L" in g.files:
        L = post.L
        assert L.shape == g[tag + "_L"].shape and np.all(np.tril(L, -1) == 0)
        assert rel(L, g[tag + "_L"]) < 1e-9
    out = m.predict(xs, ys)
    for name, v in zip(("ym", "ys2", "fm", "fs2"), out[:4]):
        assert v.shape == (xs.shape[0], 1)
        assert rel(v, g[tag + "_" + name]) < TOL, (tag, name, rel(v, g[tag + "_" + name]))
    if ys is not None:
        assert rel(out[4], g[tag + "_lp"]) < TOL
    else:
        assert out[4] is None
    return nlZ


def test_kat1_default_gpr_on_reference_fixture(golden):
    g = golden("kat_regression")
    nlZ = _check_model(pg.GPR(), g, "kat1", g["x"], g["y"], g["xs"], ys=g["ys"])
    assert abs(nlZ - 154.068967070743) < 1e-8


def test_kat2_setdata_const_mean_and_optimize(golden):
    g = golden("kat_regression")
    m = pg.GPR()
    m.setData(g["x"], g["y"])
    assert abs(m.meanfunc.hyp[0] - float(g["kat2_c"])) < 1e-15
    _check_model(m, g, "kat2", g["x"], g["y"], g["xs"])
    m = pg.GPR()
    m.setData(g["x"], g["y"])
    m.optimize(g["x"], g["y"])
    assert abs(m.nlZ - float(g["kat2_opt_nlZ"])) < 1e-5 * abs(float(g["kat2_opt_nlZ"]))
    hyp = np.array(m.meanfunc.hyp + m.covfunc.hyp + m.likfunc.hyp)
    assert np.allclose(hyp, g["kat2_opt_hyp"], rtol=1e-4, atol=1e-5), (hyp, g["kat2_opt_hyp"])


def test_kat3_other_kernels_and_linear_mean(golden):
    g = golden("kat_regression")
    x, y, xs = g["x"], g["y"], g["xs"]
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.RBFard(D=1, log_ell_list=[0.3], log_sigma=0.2))
    _check_model(m, g, "kat3_ard", x, y, xs)
    m = pg.GPR()
    m.setPrior(mean=pg.mean.Linear(D=1), kernel=pg.cov.RBF(-0.5, 0.1))
    m.setNoise(np.log(0.2))
    _check_model(m, g, "kat3_lin", x, y, xs)
    for d in (1, 3, 5, 7):
        m = pg.GPR()
        m.setPrior(kernel=pg.cov.Matern(d=d, log_ell=0.3, log_sigma=0.2))
        nlZ, post = m.getPosterior(x, y, der=False)
        ref = float(g["kat3_mat%d_nlZ" % d])
        assert abs(nlZ - ref) < 1e-8 * abs(ref), (d, nlZ, ref)
        assert rel(post.alpha, g["kat3_mat%d_alpha" % d]) < TOL
        out = m.predict(xs)
        assert rel(out[0], g["kat3_mat%d_ym" % d]) < TOL and rel(out[1], g["kat3_mat%d_ys2" % d]) < TOL


def test_matern_dnlz_by_finite_differences(golden):
    """The reference's Matern derivative is wrong: check ours against central differences."""
    g = golden("kat_regression")
    x, y = g["x"], g["y"]
    for d in (3, 7):
        base = [0.3, 0.2]
        m = pg.GPR()
        m.setPrior(kernel=pg.cov.Matern(d=d, log_ell=base[0], log_sigma=base[1]))
        nlZ, dn, _ = m.getPosterior(x, y)
        for i in range(2):
            vals = []
            for s in (+1e-5, -1e-5):
                h = list(base); h[i] += s
                mm = pg.GPR(); mm.setPrior(kernel=pg.cov.Matern(d=d, log_ell=h[0], log_sigma=h[1]))
                vals.append(mm.getPosterior(x, y, der=False)[0])
            fd = (vals[0] - vals[1]) / 2e-5
            assert abs(dn.cov[i] - fd) < 1e-5 * max(1.0, abs(fd)), (d, i, dn.cov[i], fd)


@pytest.mark.parametrize("N", [256, 1000, 2048])
def test_c2_family_synthetic(golden, N):
    g = golden("synthetic")
    X, y = go.synth_regression(N, 8)
    Xs = np.random.default_rng(1).standard_normal((300, 8))
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.RBF(np.log(2.0), 0.0))
    _check_model(m, g, "c2_%d" % N, X, y, Xs)
    assert rel(np.diag(m.posterior.L), g["c2_%d_Ldiag" % N]) < 1e-9


def test_c1_demo_shape_and_c3_ard(golden):
    g = golden("synthetic")
    m = pg.GPR()
    m.setData(g["c1_x"], g["c1_y"])
    nlZ = _check_model(m, g, "c1", g["c1_x"], g["c1_y"], g["c1_x"][:50] + 0.1)
    assert abs(nlZ - (-379.2688218267)) < 1e-7
    X, y = go.synth_regression(1024, 32)
    Xs = np.random.default_rng(1).standard_normal((200, 32))
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.RBFard(D=32, log_ell_list=[np.log(3.0)] * 32, log_sigma=0.0))
    _check_model(m, g, "c3_1024", X, y, Xs)


def test_matern_synthetic(golden):
    g = golden("synthetic")
    X, y = go.synth_regression(700, 5)
    Xs = np.random.default_rng(1).standard_normal((100, 5))
    for d in (1, 3, 5, 7):
        m = pg.GPR()
        m.setPrior(kernel=pg.cov.Matern(d=d, log_ell=np.log(1.5), log_sigma=0.1))
        nlZ, post = m.getPosterior(X, y, der=False)
        ref = float(g["mat%d_700_nlZ" % d])
        assert abs(nlZ - ref) < 1e-8 * abs(ref)
        assert rel(post.alpha, g["mat%d_700_alpha" % d]) < TOL
        out = m.predict(Xs)
        assert rel(out[0], g["mat%d_700_ym" % d]) < TOL and rel(out[1], g["mat%d_700_ys2" % d]) < TOL


def test_big_sizes_against_frozen_reference_runs(golden):
    """C2 at N = 4096 / 8192 / 16384 (the headline size) and C3 at 4096 against the reference's own run."""
    g = golden("synthetic_big")
    for N in (4096, 8192, 16384):
        X, y = go.synth_regression(N, 8)
        m = pg.GPR()
        m.setPrior(kernel=pg.cov.RBF(np.log(2.0), 0.0))
        nlZ, post = m.getPosterior(X, y, der=False)
        ref = float(g["c2_%d_nlZ" % N])
        assert abs(nlZ - ref) < 1e-9 * abs(ref), (N, nlZ, ref)
        assert rel(post.alpha, g["c2_%d_alpha" % N]) < TOL, (N, rel(post.alpha, g["c2_%d_alpha" % N]))
        Xs = np.random.default_rng(1).standard_normal((64, 8))
        out = m.predict(Xs)
        assert rel(out[0], g["c2_%d_ym" % N]) < TOL and rel(out[1], g["c2_%d_ys2" % N]) < TOL
        if N == 4096:
            assert rel(np.diag(post.L), g["c2_%d_Ldiag" % N]) < 1e-9
        # size-independent property: residual of the linear system with a freshly built K
        if N == 4096:
            K = m.covfunc.getCovMatrix(x=X, mode='train')
            r = K @ post.alpha + 0.01 * post.alpha - y
            assert np.linalg.norm(r) / np.linalg.norm(y) < 1e-9
    assert abs(float(g["c2_16384_nlZ"]) - 60824.3036486822) < 1e-6
    X, y = go.synth_regression(4096, 32)
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.RBFard(D=32, log_ell_list=[np.log(3.0)] * 32, log_sigma=0.0))
    nlZ, post = m.getPosterior(X, y, der=False)
    assert abs(nlZ - float(g["c3_4096_nlZ"])) < 1e-9 * abs(float(g["c3_4096_nlZ"]))
    assert rel(post.alpha, g["c3_4096_alpha"]) < TOL


def test_derivatives_at_4096_by_directional_finite_difference():
    X, y = go.synth_regression(4096, 8)
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.RBF(np.log(2.0), 0.0))
    nlZ, dn, _ = m.getPosterior(X, y)
    grad = np.array(dn.cov + dn.lik)
    base = np.array([np.log(2.0), 0.0, np.log(0.1)])
    v = np.array([0.6, -0.3, 0.5])
    vals = []
    for s in (1e-5, -1e-5):
        h = base + s * v
        mm = pg.GPR(); mm.setPrior(kernel=pg.cov.RBF(h[0], h[1])); mm.setNoise(h[2])
        vals.append(mm.getPosterior(X, y, der=False)[0])
    fd = (vals[0] - vals[1]) / 2e-5
    assert abs(grad @ v - fd) < 1e-5 * abs(fd), (grad @ v, fd)


def test_housing_published_optimum(golden):
    """The only number the reference publishes for this path: optimised nlZ 214.46 (demoHousing.rst:30)."""
    g = golden("housing")
    m = pg.GPR()
    nlZ, dn, post = m.getPosterior(g["x"], g["y"])
    assert abs(nlZ - float(g["default_nlZ"])) < 1e-9 * abs(nlZ)
    assert rel(dn.cov, g["default_dcov"]) < TOL and rel(dn.lik, g["default_dlik"]) < TOL
    m = pg.GPR()
    m.optimize(g["x"], g["y"])
    assert abs(m.nlZ - 214.46) < 5e-3
    assert abs(m.nlZ - float(g["opt_nlZ"])) < 1e-5 * abs(m.nlZ)


def test_not_positive_definite_is_a_python_exception():
    """Optimizers treat a failed Cholesky as a failed trial (Core/opt.py:295-298): it must raise, not abort."""
    x = np.linspace(0, 1, 50).reshape(-1, 1)
    bad = x.copy()
    bad[17, 0] = np.nan                     # a NaN row makes pivot 18 non-positive: info = 18
    m = pg.GPR()
    with pytest.raises(np.linalg.LinAlgError):
        m.getPosterior(bad, np.sin(x))
    nlZ, _ = m.getPosterior(x, np.sin(x), der=False)   # the same handle keeps working
    assert np.isfinite(nlZ)
    m2 = pg.GPR()                           # the handle must still work afterwards
    nlZ, _ = m2.getPosterior(x, np.sin(x), der=False)
    assert np.isfinite(nlZ)


def test_composite_kernel_generic_path(golden):
    g = golden("kat_regression")
    x, y, xs = g["x"], g["y"], g["xs"]
    k = pg.cov.RBF(0.0, 0.0) + pg.cov.Matern(0.3, 3, -1.0)
    m = pg.GPR()
    m.setPrior(kernel=k)
    nlZ, dn, post = m.getPosterior(x, y)
    K = (go.cov_matrix(("rbf", [0., 0.]), x=x, mode="train") + go.cov_matrix(("matern", [0.3, -1.0], 3), x=x, mode="train"))
    sn2 = 0.01
    import scipy.linalg as sla
    c = sla.cho_factor(K + sn2 * np.eye(len(x)))
    alpha = sla.cho_solve(c, y)
    ref = (y.T @ alpha / 2 + np.log(np.diag(c[0])).sum() + len(x) * np.log(2 * np.pi) / 2)[0, 0]
    assert abs(nlZ - ref) < 1e-8 * abs(ref)
    assert len(dn.cov) == 4
    ym = m.predict(xs)[0]
    ks = (go.cov_matrix(("rbf", [0., 0.]), x=x, z=xs, mode="cross") + go.cov_matrix(("matern", [0.3, -1.0], 3), x=x, z=xs, mode="cross"))
    assert rel(ym, ks.T @ alpha) < 1e-7


def test_lazy_factor_survives_the_next_evaluation(golden):
    g = golden("kat_regression")
    m = pg.GPR()
    nlZ, post = m.getPosterior(g["x"], g["y"], der=False)
    m.covfunc.hyp = [0.5, 0.1]
    m.getPosterior(g["x"], g["y"], der=False)           # overwrites the resident factor
    assert rel(post.L, g["kat1_L"]) < 1e-9              # ... but the earlier posterior still has its own L


def test_reference_shape_and_type_contract():
    """The assertions of the reference's own Testing/unit_test_inf.py:30-56 and unit_test_model.py:43-56."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((20, 2)); y = rng.standard_normal((20, 1)); z = rng.standard_normal((10, 2))
    post, nlZ, dnlZ = pg.inf.Exact().evaluate(pg.mean.Zero(), pg.cov.RBF(), pg.lik.Gauss(), x, y, nargout=3)
    assert post.alpha.shape[0] == 20 and post.L.shape == (20, 20) and post.sW.shape == (20, 1)
    assert type(nlZ) is np.float64 and all(type(v) is np.float64 for v in dnlZ.cov + dnlZ.lik)
    k = pg.cov.RBF()
    assert k.getCovMatrix(x=x, mode='train').shape == (20, 20)
    assert k.getCovMatrix(x=x, z=z, mode='cross').shape == (20, 10)
    assert k.getCovMatrix(z=z, mode='self_test').shape == (10, 1)
    assert np.min(np.linalg.eigvalsh(k.getCovMatrix(x=x, mode='train'))) > -1e-9
    m = pg.GPR()
    m.setOptimizer("Minimize", num_restarts=3)
    np.random.seed(0)
    m.optimize(x, y)
    ym, ys2, fm, fs2, lp = m.predict(z)
    assert ym.shape == ys2.shape == fm.shape == fs2.shape == (10, 1) and lp is None


def test_int8_tensor_core_and_dmma_updates_agree(monkeypatch):
    """The same evaluation with the trailing updates on the int8 tensor cores (default) and on fp64 DMMA
    (GPK_OZAKI=0): nlZ and alpha agree far inside the 1e-6 parity bar."""
    import math
    from pygps_b200 import _lib
    rng = np.random.default_rng(11)
    N = 6144
    X = rng.standard_normal((N, 6))
    y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(N)
    eng = _lib.Engine(0)
    eng.set_data(X)
    res = {}
    for oz in ("1", "0"):
        monkeypatch.setenv("GPK_OZAKI", oz)
        monkeypatch.setenv("GPK_POTRF_W1_MINREM", "8")     # make the three-level path run at this size
        out = eng.exact_eval(_lib.COV_RBF, 3, [math.log(1.5), 0.0], math.log(0.1), y, False)
        res[oz] = (out[0], np.array(out[1]))
    assert abs(res["1"][0] - res["0"][0]) <= 1e-11 * abs(res["0"][0])
    assert np.max(np.abs(res["1"][1] - res["0"][1])) <= 1e-8 * np.max(np.abs(res["0"][1]))


def test_int8_and_dmma_derivative_paths_agree(monkeypatch):
    """dnlZ with (K/sn2+I)^-1 = U U' on the int8 tensor cores (blocked L^-T through the stacked-operand sliced GEMM,
    U U' as a trapezoid sliced SYRK; default) and on fp64 DMMA (GPK_OZAKI_DER=0), at a size whose column blocks are
    ragged (T = 41 panels, blocks of 8), with an ARD kernel (10 hyper-parameters)."""
    import math
    from pygps_b200 import _lib
    rng = np.random.default_rng(12)
    N, D = 5200, 9
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :3].sum(1)) + 0.1 * rng.standard_normal(N)
    hyp = [math.log(2.5)] * D + [0.1]
    eng = _lib.Engine(0)
    eng.set_data(X)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GPK_OZAKI_DER", mode)
        out = eng.exact_eval(_lib.COV_RBFARD, 3, hyp, math.log(0.2), y, True)
        res[mode] = (out[0], np.array(out[2]), np.array(out[3]))
    assert res["1"][0] == res["0"][0]                       # the factorisation itself is the same code
    scale = np.max(np.abs(res["0"][1]))
    assert np.max(np.abs(res["1"][1] - res["0"][1])) <= 1e-10 * scale
    assert abs(res["1"][2][0] - res["0"][2][0]) <= 1e-10 * abs(res["0"][2][0])


@pytest.mark.parametrize("env", [{"GPK_POTRF_SPLIT": "0"}, {"GPK_POTRF_HEADL1": "0"}, {"GPK_LAZY_COV": "0"},
                                 {"GPK_POTRF_W2B": "6", "GPK_POTRF_W1": "6"}, {"GPK_TRSV_PERSIST": "0"}])
def test_schedule_variants_give_the_same_evaluation(monkeypatch, env):
    """The split panel chain, the head of the level-1 hand-over, the overlapped matrix build and the blocking widths only
    reorder independent work: nlZ and alpha agree with the default schedule to rounding."""
    import math
    from pygps_b200 import _lib
    rng = np.random.default_rng(13)
    N = 9000                                               # T = 71 panels: level-1 blocks of 9, then the small blocks
    X = rng.standard_normal((N, 5))
    y = np.cos(X.sum(1)) + 0.1 * rng.standard_normal(N)
    eng = _lib.Engine(0)
    eng.set_data(X)
    ref = eng.exact_eval(_lib.COV_RBF, 3, [math.log(1.7), 0.2], math.log(0.15), y, False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    out = eng.exact_eval(_lib.COV_RBF, 3, [math.log(1.7), 0.2], math.log(0.15), y, False)
    assert abs(out[0] - ref[0]) <= 1e-11 * abs(ref[0])
    assert np.max(np.abs(np.array(out[1]) - np.array(ref[1]))) <= 1e-8 * np.max(np.abs(np.array(ref[1])))


def test_int8_and_dmma_predict_solves_agree(monkeypatch):
    """GP.predict's multi-right-hand-side forward solve as a blocked sweep with int8 tensor-core updates (default for
    >= 1024 test points and >= 16 panels) against the fp64 DMMA sweep (GPK_OZAKI_PREDICT=0): ym and ys2."""
    X, y = go.synth_regression(3000, 8)
    Xs = np.random.default_rng(3).standard_normal((1500, 8))
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.RBF(np.log(2.0), 0.0))
    m.getPosterior(X, y, der=False)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GPK_OZAKI_PREDICT", mode)
        out[mode] = m.predict(Xs)
    for a, b in zip(out["1"][:4], out["0"][:4]):
        assert np.max(np.abs(a - b)) <= 1e-9 * max(np.max(np.abs(b)), 1e-300)
    # and against the oracle on a subset (the reference's own batch loop, Core/gp.py:402-419)
    hyp, sn = [np.log(2.0), 0.0], np.log(0.1)
    rpost = go.exact_evaluate(("zero",), ("rbf", hyp), sn, X, y, 1)
    if isinstance(rpost, tuple):
        rpost = rpost[0]
    rym, rys2 = go.predict(("zero",), ("rbf", hyp), sn, X, rpost, Xs[:200])[:2]
    assert rel(out["1"][0][:200], rym) < 1e-6 and rel(out["1"][1][:200], rys2) < 1e-6
