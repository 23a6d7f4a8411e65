"""Composite kernels and the remaining stationary kernels on the device (csrc/covprog.cu), and cov.Pre
(f4): matrices and derivative matrices against the reference's frozen outputs, the Mauna Loa model of
Demo/MaunaLoa/demo_MaunaLoa.py:65-68 end to end in ONE foreign call, precomputed kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import pygps_b200 as pg                      # noqa: E402
from oracle import gp_oracle as go           # noqa: E402
from parity_report import check              # noqa: E402
from test_oracle import PROGRAM_SPECS, BROKEN_IN_REFERENCE, MAUNA   # noqa: E402


def build(spec):
    """oracle spec tuple -> pygps_b200.cov kernel object"""
    c = pg.cov
    k, h = spec[0], spec[1]
    if k == "sum":
        return build(spec[1]) + build(spec[2])
    if k == "prod":
        return build(spec[1]) * build(spec[2])
    if k == "scale":
        return build(spec[2]) * float(spec[1])
    if k == "rbf":
        return c.RBF(h[0], h[1])
    if k == "rbfard":
        return c.RBFard(log_ell_list=list(h[:-1]), log_sigma=h[-1])
    if k == "matern":
        return c.Matern(h[0], spec[2], h[1])
    if k == "rbfunit":
        return c.RBFunit(h[0])
    if k == "rq":
        return c.RQ(h[0], h[1], h[2])
    if k == "rqard":
        return c.RQard(log_ell_list=list(h[:-2]), log_sigma=h[-2], log_alpha=h[-1])
    if k == "periodic":
        return c.Periodic(h[0], h[1], h[2])
    if k == "piecepoly":
        return c.PiecePoly(h[0], spec[2], h[1])
    if k == "gabor":
        return c.Gabor(h[0], h[1])
    if k == "noise":
        return c.Noise(h[0])
    if k == "const":
        return c.Const(h[0])
    if k == "linear":
        return c.Linear(h[0])
    if k == "poly":
        return c.Poly(h[0], spec[2], h[1])
    raise ValueError(k)


@pytest.mark.parametrize("name", sorted(PROGRAM_SPECS))
def test_program_matrices_and_derivatives_match_the_reference(golden, name):
    g = golden("cov_programs")
    spec, D = PROGRAM_SPECS[name]
    x, z = (g["x3"], g["z3"]) if D == 3 else (g["x1"], g["z1"])
    k = build(spec)
    assert np.allclose(k.hyp, g[name + "_hyp"])
    check(name + " K train", k.getCovMatrix(x=x, mode="train"), g[name + "_train"], 1e-12)
    check(name + " K cross", k.getCovMatrix(x=x, z=z, mode="cross"), g[name + "_cross"], 1e-12)
    self_ref = g[name + "_self"]
    got = k.getCovMatrix(z=z, mode="self_test")
    assert got.shape == self_ref.shape and np.allclose(got, self_ref, rtol=1e-12, atol=1e-14)
    for i in range(len(k.hyp)):
        for mode, key, kw in (("train", "dtrain", dict(x=x)), ("cross", "dcross", dict(x=x, z=z)),
                              ("self_test", "dself", dict(z=z))):
            ref = g["%s_%s%d" % (name, key, i)]
            if (name, i) in BROKEN_IN_REFERENCE:             # the true derivative (FD-checked in test_oracle)
                ref = go.cov_der_matrix(spec, mode=mode, der=i, **kw)
            got = k.getDerMatrix(mode=mode, der=i, **kw)
            assert got.shape == ref.shape
            assert np.allclose(got, ref, rtol=1e-10, atol=1e-12 * max(1.0, np.abs(ref).max())), (name, mode, i)
    with pytest.raises(Exception):
        k.getDerMatrix(x=x, mode="train", der=len(k.hyp) + 3)


def test_mauna_loa_composite_in_one_foreign_call(golden):
    """k1 + k2 + k3 + k4 (RBF, Periodic*RBF, RQ, RBF+Noise; 13 hyper-parameters): nlZ, all derivatives, alpha and the
    20-year extrapolation against the unmodified reference at 1e-6."""
    g = golden("cov_programs")
    X, Y, xs = g["mauna_x"], g["mauna_y"], g["mauna_xs"]
    m = pg.GPR()
    m.setData(X, Y)
    m.setPrior(kernel=build(MAUNA))
    assert abs(m.meanfunc.hyp[0] - float(g["mauna_c"])) < 1e-12
    nlZ, dn, post = m.getPosterior()
    st = m.inffunc._engine.stats()
    check("mauna nlZ", nlZ, float(g["mauna_nlZ"]), 1e-9)
    check("mauna dnlZ.cov (13)", dn.cov, g["mauna_dcov"])
    check("mauna dnlZ.lik", dn.lik, g["mauna_dlik"])
    check("mauna dnlZ.mean", dn.mean, g["mauna_dmean"])
    check("mauna alpha", post.alpha, g["mauna_alpha"])
    assert len(dn.cov) == 13 and all(type(v) is np.float64 for v in dn.cov)
    # one foreign call: nothing but y-m went up and alpha / scalars came back - no n x n matrix crossed PCIe
    n = X.shape[0]
    assert st["h2d_bytes"] < 64 * n and st["d2h_bytes"] < 64 * n
    out = m.predict(xs)
    check("mauna ym", out[0], g["mauna_ym"])
    check("mauna ys2", out[1], g["mauna_ys2"])
    check("mauna fs2", out[3], g["mauna_fs2"])
    # lazily fetched factor of the composite
    R = post.L
    assert R.shape == (n, n) and np.all(np.tril(R, -1) == 0)
    K = go.cov_matrix(MAUNA, x=X, mode="train")
    check("mauna post.L' post.L", R.T @ R, K / 0.01 + np.eye(n), 1e-10)


def test_composite_optimize_improves_nlz(golden):
    g = golden("cov_programs")
    X, Y = g["mauna_x"][::4], g["mauna_y"][::4]
    m = pg.GPR()
    m.setData(X, Y)
    m.setPrior(kernel=build(MAUNA))
    nlZ0 = m.getPosterior(der=False)[0]
    m.optimize(numIterations=8)
    assert m.nlZ < nlZ0 and len(m.covfunc.hyp) == 13


def test_composite_at_n4096_runs_the_int8_factorisation_and_matches_the_oracle():
    rng = np.random.default_rng(2)
    N, D = 4096, 4
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    spec = ("sum", ("prod", ("rbf", [0.5, 0.0]), ("rq", [0.8, 0.1, 0.3])), ("scale", -1.0, ("matern", [0.3, -0.5], 3)))
    m = pg.GPR()
    m.setPrior(kernel=build(spec))
    nlZ, dn, post = m.getPosterior(X, y)
    rpost, rnlZ, rdn = go.exact_evaluate(("zero",), spec, np.log(0.1), X, y, 3)
    check("composite N=4096 nlZ", nlZ, rnlZ, 1e-9)
    check("composite N=4096 dcov", dn.cov, rdn["cov"])
    check("composite N=4096 alpha", post.alpha, rpost["alpha"])
    Xs = rng.standard_normal((50, D))
    out = m.predict(Xs)
    rym, rys2 = go.predict(("zero",), spec, np.log(0.1), X, rpost, Xs)[:2]
    check("composite N=4096 ym", out[0], rym)
    check("composite N=4096 ys2", out[1], rys2)


def test_precomputed_kernel_matrices(golden):
    """cov.Pre (Core/cov.py:1429-1455): M2 uploaded once, evaluation on the device; predictions from M1."""
    g = golden("cov_programs")
    x, z, y = g["x3"], g["z3"], g["pre_y"]
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.Pre(g["pre_M1"], g["pre_M2"]))
    nlZ, post = m.getPosterior(x, y, der=False)
    check("pre nlZ", nlZ, float(g["pre_nlZ"]), 1e-10)
    check("pre alpha", post.alpha, g["pre_alpha"])
    out = m.predict(z)
    check("pre ym", out[0], g["pre_ym"])
    check("pre ys2", out[1], g["pre_ys2"])
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.Pre(g["pre_M1"], g["pre_M2"]) + pg.cov.Noise(-1.0))
    nlZ, dn, post = m.getPosterior(x, y)
    check("pre+noise nlZ", nlZ, float(g["prenoise_nlZ"]), 1e-10)
    check("pre+noise dcov", dn.cov, g["prenoise_dcov"])
    check("pre+noise dlik", dn.lik, g["prenoise_dlik"])
    check("pre+noise alpha", post.alpha, g["prenoise_alpha"])
    with pytest.raises(Exception):
        pg.cov.Pre(g["pre_M1"], g["pre_M2"]).getDerMatrix(x=x, mode="train", der=0)


def test_unsupported_kernels_fail_loudly():
    class Mine(pg.cov.Kernel):
        def __init__(self):
            self.hyp = [0.0]
            self.para = []
    x = np.random.default_rng(0).standard_normal((10, 2)); y = x[:, :1]
    m = pg.GPR()
    m.setPrior(kernel=Mine() + pg.cov.RBF())
    with pytest.raises(Exception) as ei:
        m.getPosterior(x, y)
    assert "no device implementation" in str(ei.value)


def test_program_with_a_varying_diagonal_through_the_big_block_int8_path():
    """N=7424 (58 panels: one level-1 block of 9 panels, sliced sub-block by sub-block with FIXED row scales) and a
    composite whose diagonal varies from point to point (Linear term): the scales are then read from the diagonal of the
    built matrix and the panel stream has to wait for them (potrf_device: ev_fix).  Against the CPU oracle."""
    rng = np.random.default_rng(9)
    N, D = 7424, 3
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    spec = ("sum", ("prod", ("rbf", [0.4, 0.2]), ("const", [0.1])), ("linear", [np.log(0.3)]))
    m = pg.GPR()
    m.setPrior(kernel=build(spec))
    nlZ, post = m.getPosterior(X, y, der=False)
    rpost, rnlZ = go.exact_evaluate(("zero",), spec, np.log(0.1), X, y, 2)
    check("varying-diagonal composite N=7424 nlZ", nlZ, rnlZ, 1e-9)
    check("varying-diagonal composite N=7424 alpha", post.alpha, rpost["alpha"])


def test_ragged_size_on_the_int8_path_matches_the_oracle():
    """N=5003 (np = 5120: 117 identity-padded rows) through the fixed-scale int8 factorisation, native RBF kernel."""
    rng = np.random.default_rng(10)
    N, D = 5003, 6
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    m = pg.GPR()
    m.setPrior(kernel=pg.cov.RBF(np.log(1.8), 0.1))
    nlZ, dn, post = m.getPosterior(X, y)
    rpost, rnlZ, rdn = go.exact_evaluate(("zero",), ("rbf", [np.log(1.8), 0.1]), np.log(0.1), X, y, 3)
    check("ragged N=5003 nlZ", nlZ, rnlZ, 1e-10)
    check("ragged N=5003 alpha", post.alpha, rpost["alpha"])
    check("ragged N=5003 dcov", dn.cov, rdn["cov"])
    check("ragged N=5003 dlik", dn.lik, rdn["lik"])
