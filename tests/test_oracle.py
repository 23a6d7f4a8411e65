"""The CPU oracle (oracle/gp_oracle.py) pinned against outputs of the UNMODIFIED reference
frozen in tests/golden/*.npz by oracle/gen_golden.py.  No GPU needed."""
import numpy as np
import pytest

from oracle import gp_oracle as go

RT = 1e-11


def close(a, b, rtol=RT, atol=1e-12):
    np.testing.assert_allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=rtol, atol=atol)


def test_cov_matrices_all_modes(golden):
    g = golden("cov_vectors")
    x, z = g["x"], g["z"]
    specs = {"rbf": ("rbf", list(g["rbf_hyp"])), "ard": ("rbfard", list(g["ard_hyp"]))}
    for d in (1, 3, 5, 7):
        specs["mat%d" % d] = ("matern", list(g["mat%d_hyp" % d]), d)
    for name, cov in specs.items():
        close(go.cov_matrix(cov, x=x, mode="train"), g[name + "_train"])
        close(go.cov_matrix(cov, x=x, z=z, mode="cross"), g[name + "_cross"])
        close(go.cov_matrix(cov, z=z, mode="self_test"), g[name + "_self"])
    for name in ("rbf", "ard"):
        cov = specs[name]
        for i in range(len(cov[1])):
            close(go.cov_der_matrix(cov, x=x, mode="train", der=i), g["%s_dtrain%d" % (name, i)])
            close(go.cov_der_matrix(cov, x=x, z=z, mode="cross", der=i), g["%s_dcross%d" % (name, i)])
    dk, kuu, ku = go.fitc_cov_matrix(specs["rbf"], g["u"], x=x, mode="train")
    close(dk, g["fitc_diag"]); close(kuu, g["fitc_kuu"]); close(ku, g["fitc_ku"])
    close(go.fitc_cov_matrix(specs["rbf"], g["u"], x=x, z=z, mode="cross"), g["fitc_cross"])


def test_matern_derivative_is_the_true_gradient(golden):
    """The reference's Matern.getDerMatrix is wrong; ours is checked by finite differences."""
    g = golden("cov_vectors")
    x = g["x"]
    for d in (1, 3, 5, 7):
        hyp = list(g["mat%d_hyp" % d])
        for i in range(2):
            hp, hm = list(hyp), list(hyp)
            hp[i] += 1e-6; hm[i] -= 1e-6
            fd = (go.cov_matrix(("matern", hp, d), x=x, mode="train")
                  - go.cov_matrix(("matern", hm, d), x=x, mode="train")) / 2e-6
            np.testing.assert_allclose(go.cov_der_matrix(("matern", hyp, d), x=x, mode="train", der=i), fd,
                                       rtol=1e-5, atol=1e-8)


def test_jitchol_solve_chol(golden):
    g = golden("cov_vectors")
    L = go.jitchol(g["chol_A"])
    close(L, g["chol_L"])
    close(go.solve_chol(L.T, g["chol_B"]), g["chol_X"], rtol=1e-9)
    bad = g["chol_A"].copy()
    bad[3, 3] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        go.jitchol(bad)
    with pytest.raises(Exception):
        go.solve_chol(L.T, np.zeros((3, 1)))


def _check_exact(g, tag, mean, cov, log_sn, x, y, xs, ys=None, der=True):
    post, nlZ, dn = go.exact_evaluate(mean, cov, log_sn, x, y, 3)
    close(nlZ, g[tag + "_nlZ"])
    close(post["alpha"], g[tag + "_alpha"], rtol=1e-8)
    close(post["sW"][0, 0], g[tag + "_sW0"])
    if tag + "_L" in g.files:
        close(post["L"], g[tag + "_L"], rtol=1e-9)
    if der:
        close(dn["cov"], g[tag + "_dcov"], rtol=1e-7)
        close(dn["lik"], g[tag + "_dlik"], rtol=1e-7)
        close(dn["mean"], g[tag + "_dmean"], rtol=1e-7)
    out = go.predict(mean, cov, log_sn, x, post, xs, ys)
    for name, v in zip(("ym", "ys2", "fm", "fs2"), out[:4]):
        close(v, g[tag + "_" + name], rtol=1e-7, atol=1e-10)
    if ys is not None:
        close(out[4], g[tag + "_lp"], rtol=1e-7)


def test_exact_kats_on_reference_fixture(golden):
    g = golden("kat_regression")
    x, y, xs = g["x"], g["y"], g["xs"]
    _check_exact(g, "kat1", ("zero",), ("rbf", [0., 0.]), np.log(0.1), x, y, xs, ys=g["ys"])
    assert abs(float(g["kat1_nlZ"]) - 154.068967070743) < 1e-9
    _check_exact(g, "kat2", ("const", float(g["kat2_c"])), ("rbf", [0., 0.]), np.log(0.1), x, y, xs)
    _check_exact(g, "kat3_ard", ("zero",), ("rbfard", [0.3, 0.2]), np.log(0.1), x, y, xs)
    _check_exact(g, "kat3_lin", ("linear", [0.5]), ("rbf", [-0.5, 0.1]), np.log(0.2), x, y, xs)
    for d in (1, 3, 5, 7):
        post, nlZ = go.exact_evaluate(("zero",), ("matern", [0.3, 0.2], d), np.log(0.1), x, y, 2)
        close(nlZ, g["kat3_mat%d_nlZ" % d])
        close(post["alpha"], g["kat3_mat%d_alpha" % d], rtol=1e-8)
        out = go.predict(("zero",), ("matern", [0.3, 0.2], d), np.log(0.1), x, post, xs)
        close(out[0], g["kat3_mat%d_ym" % d], rtol=1e-7)
        close(out[1], g["kat3_mat%d_ys2" % d], rtol=1e-7)


def test_fitc_kats(golden):
    g = golden("kat_regression")
    x, y, xs = g["x"], g["y"], g["xs"]
    for tag, mean, u in (("kat4", ("zero",), g["u"]), ("kat4b", ("const", float(g["kat4b_c"])), g["kat4b_u"])):
        cov = ("rbf", [0., 0.])
        post, nlZ, dn = go.fitc_evaluate(mean, cov, u, np.log(0.1), x, y, 3)
        close(nlZ, g[tag + "_nlZ"], rtol=1e-9)
        close(post["alpha"], g[tag + "_alpha"], rtol=1e-6)
        close(post["L"], g[tag + "_L"], rtol=1e-5, atol=1e-6)
        close(dn["cov"], g[tag + "_dcov"], rtol=1e-6)
        close(dn["lik"], g[tag + "_dlik"], rtol=1e-6)
        close(dn["mean"], g[tag + "_dmean"], rtol=1e-6)
        out = go.predict(mean, cov, np.log(0.1), x, post, xs, xu=u)
        close(out[0], g[tag + "_ym"], rtol=1e-6)
        close(out[1], g[tag + "_ys2"], rtol=1e-6)
    assert abs(float(g["kat4_nlZ"]) - 179.399011418159) < 1e-8


def test_synthetic_baseline_configs(golden):
    g = golden("synthetic")
    for N in (256, 1000):
        X, y = go.synth_regression(N, 8)
        Xs = np.random.default_rng(1).standard_normal((300, 8))
        _check_exact(g, "c2_%d" % N, ("zero",), ("rbf", [np.log(2.0), 0.0]), np.log(0.1), X, y, Xs)
    X, y = go.synth_regression(2048, 8)
    assert abs(go.exact_evaluate_fair(("rbf", [np.log(2.0), 0.0]), np.log(0.1), X, y) - float(g["c2_2048_nlZ"])) \
        < 1e-7 * abs(float(g["c2_2048_nlZ"]))
    _check_exact(g, "c1", ("const", float(g["c1_c"])), ("rbf", [0., 0.]), np.log(0.1), g["c1_x"], g["c1_y"],
                 g["c1_x"][:50] + 0.1)
    assert abs(float(g["c1_nlZ"]) - (-379.2688218267)) < 1e-8
    X, y = go.synth_regression(1024, 32)
    Xs = np.random.default_rng(1).standard_normal((200, 32))
    _check_exact(g, "c3_1024", ("zero",), ("rbfard", [np.log(3.0)] * 32 + [0.0]), np.log(0.1), X, y, Xs)


def test_fitc_synthetic(golden):
    g = golden("synthetic")
    N, M = 2000, 100
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, 8))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    U = rng.standard_normal((M, 8))
    post, nlZ, dn = go.fitc_evaluate(("zero",), ("rbf", [np.log(2.0), 0.0]), U, np.log(0.1), X, y, 3)
    tag = "c4_%d_%d" % (N, M)
    close(nlZ, g[tag + "_nlZ"], rtol=1e-9)
    close(dn["cov"], g[tag + "_dcov"], rtol=1e-5)
    close(dn["lik"], g[tag + "_dlik"], rtol=1e-5)


def test_fitc_larger_pin_size(golden):
    """SURVEY 8(c)/(d) larger config-4 size (32768, 512) against the golden frozen from the unmodified reference."""
    g = golden("synthetic_c4big")
    N, M = 32768, 512
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, 8))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    U = rng.standard_normal((M, 8))
    post, nlZ = go.fitc_evaluate(("zero",), ("rbf", [np.log(2.0), 0.0]), U, np.log(0.1), X, y, 2)
    close(nlZ, g["c4_%d_%d_nlZ" % (N, M)], rtol=1e-9)
    close(post["alpha"], g["c4_%d_%d_alpha" % (N, M)], rtol=1e-5, atol=1e-6 * np.abs(g["c4_%d_%d_alpha" % (N, M)]).max())


def test_housing_published_value(golden):
    """doc/source/demoHousing.rst:30 publishes the optimised nlZ 214.46; the frozen run reproduces it."""
    g = golden("housing")
    assert abs(float(g["opt_nlZ"]) - 214.46) < 5e-3
    post, nlZ, dn = go.exact_evaluate(("zero",), ("rbf", [0., 0.]), np.log(0.1), g["x"], g["y"], 3)
    close(nlZ, g["default_nlZ"])
    close(dn["cov"], g["default_dcov"], rtol=1e-7)
    close(dn["lik"], g["default_dlik"], rtol=1e-7)
    h = g["opt_hyp"]
    _, nlZo = go.exact_evaluate(("zero",), ("rbf", [h[0], h[1]]), h[2], g["x"], g["y"], 2)
    close(nlZo, g["opt_nlZ"], rtol=1e-9)


def test_ep_classification_against_reference(golden):
    """inf.EP + lik.Erf (config 5 family): KAT5 on Demo/Classification and synthetic sign labels."""
    g = golden("classification")
    x, y, xs = g["kat5_x"], g["kat5_y"], g["kat5_xs"]
    post, nlZ, dn, extra = go.ep_evaluate(("zero",), ("rbf", [0., 0.]), x, y, nargout=3)
    close(nlZ, g["kat5_nlZ"], rtol=1e-9)
    assert abs(float(g["kat5_nlZ"]) - 50.454379530956) < 1e-8
    close(dn["cov"], g["kat5_dcov"], rtol=1e-6)
    close(post["alpha"], g["kat5_alpha"], rtol=1e-7, atol=1e-10)
    close(post["sW"], g["kat5_sW"], rtol=1e-7)
    close(post["L"], g["kat5_L"], rtol=1e-7, atol=1e-10)
    close(extra["ttau"], g["kat5_ttau"], rtol=1e-7)
    out = go.predict_class(("zero",), ("rbf", [0., 0.]), x, post, xs, np.ones((xs.shape[0], 1)))
    for name, v in zip(("ym", "ys2", "fm", "fs2", "lp"), out[:5]):
        close(v, g["kat5_" + name], rtol=1e-6, atol=1e-9)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((200, 16))
    lab = np.sign(X[:, :1] + 0.5 * X[:, 1:2] + 0.3 * rng.standard_normal((200, 1)))
    lab[lab == 0] = 1
    post, nlZ, dn, extra = go.ep_evaluate(("zero",), ("rbf", [np.log(4.0), 0.0]), X, lab, nargout=3)
    close(nlZ, g["c5_200_nlZ"], rtol=1e-9)
    close(dn["cov"], g["c5_200_dcov"], rtol=1e-6)
    close(post["alpha"], g["c5_200_alpha"], rtol=1e-6, atol=1e-10)


def test_erf_special_functions_cover_all_branches():
    z = np.array([-8.0, -6.2, -6.0, -5.8, -5.5, -5.2, -5.0, -3.0, 0.0, 2.0, 7.0])
    p = (1 + __import__("scipy.special", fromlist=["erf"]).erf(z / np.sqrt(2))) / 2
    lp = go.erf_logphi(z, p)
    assert np.all(np.isfinite(lp)) and np.all(np.diff(lp) > 0)
    from scipy.stats import norm
    np.testing.assert_allclose(lp, norm.logcdf(z), rtol=2e-3, atol=1e-3)
    n_p = go.erf_gau_over_cum_gauss(z, p)
    np.testing.assert_allclose(n_p[z > -5], norm.pdf(z[z > -5]) / norm.cdf(z[z > -5]), rtol=1e-10)
    assert np.all(np.diff(n_p) < 0)


PROGRAM_SPECS = {
    "rbfunit": (("rbfunit", [0.3]), 3),
    "rq": (("rq", [0.2, -0.1, 0.4]), 3),
    "rqard": (("rqard", [0.1, -0.2, 0.3, 0.2, -0.3]), 3),
    "periodic": (("periodic", [0.2, 0.5, 0.1]), 1),
    "piecepoly0": (("piecepoly", [0.9, 0.1], 0), 3),
    "piecepoly1": (("piecepoly", [0.9, 0.1], 1), 3),
    "piecepoly2": (("piecepoly", [0.9, 0.1], 2), 3),
    "piecepoly3": (("piecepoly", [1.2, -0.2], 3), 3),
    "gabor": (("gabor", [0.4, 0.3]), 3),
    "noise": (("noise", [-0.7]), 3),
    "const": (("const", [0.3]), 3),
    "linear": (("linear", [-0.4]), 3),
    "poly": (("poly", [0.2, -0.3], 3), 3),
    "comp3": (("sum", ("sum", ("sum", ("prod", ("rbf", [0.3, 0.1]), ("matern", [0.5, -0.2], 5)),
                               ("scale", 1.5, ("rq", [0.2, -0.1, 0.4]))), ("noise", [-1.0])),
               ("prod", ("const", [-0.5]), ("linear", [-1.0]))), 3),
    "comp1": (("sum", ("sum", ("sum", ("rbf", [1.0, 0.5]), ("prod", ("periodic", [0.2, 0.5, 0.1]), ("rbf", [1.5, -0.2]))),
                       ("rq", [0.1, -0.4, -0.2])), ("sum", ("rbf", [-1.0, -1.0]), ("noise", [-1.5]))), 1),
}
# derivative matrices the reference gets structurally wrong (checked by finite differences instead):
#   Matern, BOTH derivatives (Core/cov.py:1173-1177: the distance is overwritten by K before it is used);
#   RQard length scales (:1413-1420: identically zero in train mode, scaled by ell instead of 1/ell in cross mode)
BROKEN_IN_REFERENCE = {("rqard", 0), ("rqard", 1), ("rqard", 2), ("comp3", 2), ("comp3", 3)}


def test_program_kernels_and_composites_against_reference(golden):
    g = golden("cov_programs")
    for name, (spec, D) in PROGRAM_SPECS.items():
        x, z = (g["x3"], g["z3"]) if D == 3 else (g["x1"], g["z1"])
        close(go.cov_matrix(spec, x=x, mode="train"), g[name + "_train"], rtol=1e-12, atol=1e-14)
        close(go.cov_matrix(spec, x=x, z=z, mode="cross"), g[name + "_cross"], rtol=1e-12, atol=1e-14)
        close(go.cov_matrix(spec, z=z, mode="self_test"), g[name + "_self"], rtol=1e-12, atol=1e-14)
        assert go.cov_nhyp(spec, D) == len(g[name + "_hyp"])
        for i in range(go.cov_nhyp(spec, D)):
            if (name, i) in BROKEN_IN_REFERENCE:
                continue
            close(go.cov_der_matrix(spec, x=x, mode="train", der=i), g["%s_dtrain%d" % (name, i)], rtol=1e-11, atol=1e-13)
            close(go.cov_der_matrix(spec, x=x, z=z, mode="cross", der=i), g["%s_dcross%d" % (name, i)], rtol=1e-11, atol=1e-13)
            close(go.cov_der_matrix(spec, z=z, mode="self_test", der=i), g["%s_dself%d" % (name, i)], rtol=1e-11, atol=1e-13)


def test_true_derivatives_where_the_reference_is_broken(golden):
    """RQard length-scale derivatives by central differences of the (reference-pinned) covariance itself."""
    g = golden("cov_programs")
    x = g["x3"]
    hyp = [0.1, -0.2, 0.3, 0.2, -0.3]
    for i in range(3):
        hp, hm = list(hyp), list(hyp)
        hp[i] += 1e-6; hm[i] -= 1e-6
        fd = (go.cov_matrix(("rqard", hp), x=x, mode="train") - go.cov_matrix(("rqard", hm), x=x, mode="train")) / 2e-6
        close(go.cov_der_matrix(("rqard", hyp), x=x, mode="train", der=i), fd, rtol=1e-6, atol=1e-8)
        assert np.all(g["rqard_dtrain%d" % i] == 0)          # what the reference returns
    # the whole gradient of the 3-d composite (it contains a Matern factor) by central differences
    spec, D = PROGRAM_SPECS["comp3"]

    def rebuild(sp, h, pos=0):
        k = sp[0]
        if k in ("sum", "prod"):
            a, pos = rebuild(sp[1], h, pos)
            b, pos = rebuild(sp[2], h, pos)
            return (k, a, b), pos
        if k == "scale":
            c = h[pos]
            a, pos = rebuild(sp[2], h, pos + 1)
            return (k, c, a), pos
        n = len(sp[1])
        return (k, list(h[pos:pos + n])) + tuple(sp[2:]), pos + n

    def flat(sp):
        k = sp[0]
        if k in ("sum", "prod"):
            return flat(sp[1]) + flat(sp[2])
        if k == "scale":
            return [sp[1]] + flat(sp[2])
        return list(sp[1])
    h0 = flat(spec)
    assert rebuild(spec, h0)[0] == spec
    for i in range(len(h0)):
        hp, hm = list(h0), list(h0)
        hp[i] += 1e-6; hm[i] -= 1e-6
        fd = (go.cov_matrix(rebuild(spec, hp)[0], x=x, mode="train") - go.cov_matrix(rebuild(spec, hm)[0], x=x, mode="train")) / 2e-6
        an = go.cov_der_matrix(spec, x=x, mode="train", der=i)
        if spec_is_reference_convention(spec, i):
            an = an / 2.0          # ScaleOfKernel / Const / Linear: the reference's derivative carries a factor 2 (exp(h), not exp(2h))
        close(an, fd, rtol=2e-6, atol=1e-8)


def spec_is_reference_convention(spec, i):
    """comp3's hyper-parameter order: rbf(0,1) matern(2,3) scale(4) rq(5,6,7) noise(8) const(9) linear(10)."""
    return i in (4, 9, 10)


MAUNA = ("sum", ("sum", ("sum", ("rbf", [np.log(67.), np.log(66.)]),
                         ("prod", ("periodic", [np.log(1.3), np.log(1.0), np.log(2.4)]), ("rbf", [np.log(90.), np.log(2.4)]))),
                 ("rq", [np.log(1.2), np.log(0.66), np.log(0.78)])),
         ("sum", ("rbf", [np.log(1.6 / 12.), np.log(0.18)]), ("noise", [np.log(0.19)])))


def test_mauna_loa_composite_against_reference(golden):
    """Demo/MaunaLoa/demo_MaunaLoa.py:65-68: k1 + k2 + k3 + k4 with 13 hyper-parameters, Const mean from setData."""
    g = golden("cov_programs")
    X, Y, xs = g["mauna_x"], g["mauna_y"], g["mauna_xs"]
    mean = ("const", float(g["mauna_c"]))
    post, nlZ, dn = go.exact_evaluate(mean, MAUNA, np.log(0.1), X, Y, 3)
    close(nlZ, g["mauna_nlZ"], rtol=1e-9)
    close(dn["cov"], g["mauna_dcov"], rtol=1e-6, atol=1e-8)
    close(dn["lik"], g["mauna_dlik"], rtol=1e-6)
    close(post["alpha"], g["mauna_alpha"], rtol=1e-6, atol=1e-8 * np.abs(g["mauna_alpha"]).max())
    ym, ys2 = go.predict(mean, MAUNA, np.log(0.1), X, post, xs)[:2]
    close(ym, g["mauna_ym"], rtol=1e-8)
    close(ys2, g["mauna_ys2"], rtol=1e-6)
