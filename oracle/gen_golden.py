#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py [--big]

The reference (marionmari/pyGPs @ 792f3c6) is imported from /root/reference
through the three stub modules in oracle/shim/ (past.utils, past.builtins,
matplotlib.pyplot - SURVEY 8(c)).  Every array written here is an output of the
reference's own classes (pyGPs.GPR / GPR_FITC / cov.* / inf.*); inputs are stored
next to the outputs when small, otherwise regenerated from the stated seed.
`--big` adds the N=8192/16384 C2 runs (minutes of CPU, ~7 GiB).
"""
import os
import sys
import logging
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYGPS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
if not hasattr(np, "float"):
    # cov.Noise's cross mode (Core/cov.py:1277,1296) uses the alias np.float, removed in numpy 1.24: restore the alias
    # (same object as the builtin it always was) so the unmodified reference runs on this container's numpy 2.3
    np.float = float
import pyGPs  # noqa: E402

logging.disable(logging.WARNING)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def synth(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    return X, y


def flat_dnlz(d):
    return (np.array(d.mean, dtype=float), np.array(d.cov, dtype=float),
            np.array(d.lik, dtype=float))


def run_model(model, x, y, xs, tag, store, keep_L=True, ys=None):
    nlZ, dn, post = model.getPosterior(x, y, der=True)
    dm, dc, dl = flat_dnlz(dn)
    store[tag + "_nlZ"] = np.float64(nlZ)
    store[tag + "_dmean"], store[tag + "_dcov"], store[tag + "_dlik"] = dm, dc, dl
    store[tag + "_alpha"] = post.alpha
    store[tag + "_sW0"] = np.float64(post.sW[0, 0])
    if keep_L:
        store[tag + "_L"] = post.L
    if xs is not None:
        out = model.predict(xs, ys)
        for name, v in zip(("ym", "ys2", "fm", "fs2"), out[:4]):
            store[tag + "_" + name] = v
        if ys is not None:
            store[tag + "_lp"] = out[4]


def kat_regression():
    """KAT1-KAT4 on the reference's shipped fixture Demo/Regression/regression_data.npz."""
    d = np.load(os.path.join(REF, "pyGPs/Demo/Regression/regression_data.npz"))
    x, y, xs = d["x"], d["y"], d["xstar"]
    s = {"x": x, "y": y, "xs": xs}
    ys = np.sin(xs) + 0.3
    s["ys"] = ys

    m = pyGPs.GPR()                                   # KAT1: Zero mean, RBF(0,0), Gauss(log .1)
    run_model(m, x, y, xs, "kat1", s, ys=ys)

    m = pyGPs.GPR()                                   # KAT2: setData -> mean.Const(mean(y))
    m.setData(x, y)
    s["kat2_c"] = np.float64(m.meanfunc.hyp[0])
    run_model(m, x, y, xs, "kat2", s)
    m = pyGPs.GPR()
    m.setData(x, y)
    m.optimize(x, y)
    s["kat2_opt_nlZ"] = np.float64(m.nlZ)
    s["kat2_opt_hyp"] = np.array(m.meanfunc.hyp + m.covfunc.hyp + m.likfunc.hyp)

    m = pyGPs.GPR()                                   # KAT3: other kernels
    m.setPrior(kernel=pyGPs.cov.RBFard(D=1, log_ell_list=[0.3], log_sigma=0.2))
    run_model(m, x, y, xs, "kat3_ard", s)
    for dd in (1, 3, 5, 7):
        m = pyGPs.GPR()
        m.setPrior(kernel=pyGPs.cov.Matern(d=dd, log_ell=0.3, log_sigma=0.2))
        nlZ, post = m.getPosterior(x, y, der=False)   # Matern dnlZ is buggy in the reference
        s["kat3_mat%d_nlZ" % dd] = np.float64(nlZ)
        s["kat3_mat%d_alpha" % dd] = post.alpha
        out = m.predict(xs)
        s["kat3_mat%d_ym" % dd], s["kat3_mat%d_ys2" % dd] = out[0], out[1]
    m = pyGPs.GPR()                                   # Linear mean (Core/mean.py:339)
    m.setPrior(mean=pyGPs.mean.Linear(D=1), kernel=pyGPs.cov.RBF(-0.5, 0.1))
    m.setNoise(np.log(0.2))
    run_model(m, x, y, xs, "kat3_lin", s)

    u = np.array([[-1.], [-.8], [-.5], [.3], [1.]])   # KAT4: Testing/unit_test_model.py:33
    s["u"] = u
    m = pyGPs.GPR_FITC()
    m.setPrior(kernel=pyGPs.cov.RBF(), inducing_points=u)
    run_model(m, x, y, xs, "kat4", s)
    m = pyGPs.GPR_FITC()
    m.setData(x, y)                                   # default grid of 5 inducing points + Const mean
    s["kat4b_u"] = m.u
    s["kat4b_c"] = np.float64(m.meanfunc.hyp[0])
    run_model(m, x, y, xs, "kat4b", s)
    np.savez_compressed(os.path.join(OUT, "kat_regression.npz"), **s)
    print("kat1 nlZ", s["kat1_nlZ"], "kat4 nlZ", s["kat4_nlZ"], "kat2 opt", s["kat2_opt_nlZ"])


def cov_vectors():
    """Kernel matrices in all three modes + derivatives; inputs like Testing/unit_test_cov.py:24-30
    but with a spread of length scales so K is not the identity."""
    rng = np.random.RandomState(0)
    x = rng.normal(0, 2.0, (20, 3))
    z = rng.normal(0, 2.0, (10, 3))
    s = {"x": x, "z": z}
    kernels = {
        "rbf": pyGPs.cov.RBF(0.7, -0.3),
        "ard": pyGPs.cov.RBFard(log_ell_list=[0.5, -0.2, 1.1], log_sigma=0.4),
        "mat1": pyGPs.cov.Matern(0.6, 1, 0.2),
        "mat3": pyGPs.cov.Matern(0.6, 3, 0.2),
        "mat5": pyGPs.cov.Matern(0.6, 5, 0.2),
        "mat7": pyGPs.cov.Matern(0.6, 7, 0.2),
    }
    for name, k in kernels.items():
        s[name + "_hyp"] = np.array(k.hyp, dtype=float)
        s[name + "_train"] = k.getCovMatrix(x=x, mode="train")
        s[name + "_cross"] = k.getCovMatrix(x=x, z=z, mode="cross")
        s[name + "_self"] = k.getCovMatrix(z=z, mode="self_test")
        if not name.startswith("mat"):
            for i in range(len(k.hyp)):
                s["%s_dtrain%d" % (name, i)] = k.getDerMatrix(x=x, mode="train", der=i)
                s["%s_dcross%d" % (name, i)] = k.getDerMatrix(x=x, z=z, mode="cross", der=i)
    u = rng.normal(0, 2.0, (5, 3))
    s["u"] = u
    f = pyGPs.cov.RBF(0.7, -0.3).fitc(u)
    dk, kuu, ku = f.getCovMatrix(x=x, mode="train")
    s["fitc_diag"], s["fitc_kuu"], s["fitc_ku"] = dk, kuu, ku
    s["fitc_cross"] = f.getCovMatrix(x=x, z=z, mode="cross")
    A = kernels["rbf"].getCovMatrix(x=x, mode="train") + 0.01 * np.eye(20)
    L = pyGPs.Core.tools.jitchol(A)
    s["chol_A"], s["chol_L"] = A, L
    B = rng.normal(size=(20, 3))
    s["chol_B"] = B
    s["chol_X"] = pyGPs.Core.tools.solve_chol(L.T, B)
    np.savez_compressed(os.path.join(OUT, "cov_vectors.npz"), **s)


def synthetic(big=False):
    """BASELINE configs at oracle-feasible sizes; inputs come from the seed."""
    s = {}
    # C2 family: RBF(log 2, 0), Gauss(log 0.1), Zero mean, D=8
    for N in (256, 1000, 2048):
        X, y = synth(N, 8)
        Xs = np.random.default_rng(1).standard_normal((300, 8))
        m = pyGPs.GPR()
        m.setPrior(kernel=pyGPs.cov.RBF(np.log(2.0), 0.0))
        run_model(m, X, y, Xs, "c2_%d" % N, s, keep_L=False)
        s["c2_%d_Ldiag" % N] = np.diag(m.posterior.L).copy()
        print("c2", N, s["c2_%d_nlZ" % N])
    # C1: N=512, D=2, default GPR via setData (Const mean), np.random.seed(0) recipe of BASELINE.md
    np.random.seed(0)
    x1 = np.random.randn(512, 2)
    y1 = np.sin(x1[:, :1]) + 0.1 * np.random.randn(512, 1)
    s["c1_x"], s["c1_y"] = x1, y1
    m = pyGPs.GPR()
    m.setData(x1, y1)
    s["c1_c"] = np.float64(m.meanfunc.hyp[0])
    run_model(m, x1, y1, x1[:50] + 0.1, "c1", s, keep_L=False)
    print("c1", s["c1_nlZ"])
    # C3 family: RBFard D=32, ell=3
    for N in (1024,):
        X, y = synth(N, 32)
        m = pyGPs.GPR()
        m.setPrior(kernel=pyGPs.cov.RBFard(D=32, log_ell_list=[np.log(3.0)] * 32, log_sigma=0.0))
        Xs = np.random.default_rng(1).standard_normal((200, 32))
        run_model(m, X, y, Xs, "c3_%d" % N, s, keep_L=False)
        print("c3", N, s["c3_%d_nlZ" % N])
    # Matern on synthetic
    X, y = synth(700, 5)
    Xs = np.random.default_rng(1).standard_normal((100, 5))
    for dd in (1, 3, 5, 7):
        m = pyGPs.GPR()
        m.setPrior(kernel=pyGPs.cov.Matern(d=dd, log_ell=np.log(1.5), log_sigma=0.1))
        nlZ, post = m.getPosterior(X, y, der=False)
        s["mat%d_700_nlZ" % dd] = np.float64(nlZ)
        s["mat%d_700_alpha" % dd] = post.alpha
        out = m.predict(Xs)
        s["mat%d_700_ym" % dd], s["mat%d_700_ys2" % dd] = out[0], out[1]
    # C4 family: GPR_FITC, RBF(log 2, 0), D=8 (fp64 - the reference has no fp32 path)
    for N, M in ((4096, 256), (2000, 100)):
        rng = np.random.default_rng(0)
        X = rng.standard_normal((N, 8))
        y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
        U = rng.standard_normal((M, 8))
        Xs = np.random.default_rng(1).standard_normal((300, 8))
        m = pyGPs.GPR_FITC()
        m.setPrior(kernel=pyGPs.cov.RBF(np.log(2.0), 0.0), inducing_points=U)
        run_model(m, X, y, Xs, "c4_%d_%d" % (N, M), s, keep_L=(M <= 100))
        print("c4", N, M, s["c4_%d_%d_nlZ" % (N, M)])
    np.savez_compressed(os.path.join(OUT, "synthetic.npz"), **s)

    if big:
        b = {}
        for N in (4096, 8192, 16384):
            X, y = synth(N, 8)
            m = pyGPs.GPR()
            m.setPrior(kernel=pyGPs.cov.RBF(np.log(2.0), 0.0))
            nlZ, post = m.getPosterior(X, y, der=False)
            b["c2_%d_nlZ" % N] = np.float64(nlZ)
            b["c2_%d_alpha" % N] = post.alpha
            b["c2_%d_Ldiag" % N] = np.diag(post.L).copy()
            Xs = np.random.default_rng(1).standard_normal((64, 8))
            out = m.predict(Xs)
            b["c2_%d_ym" % N], b["c2_%d_ys2" % N] = out[0], out[1]
            print("c2 big", N, repr(nlZ), flush=True)
            np.savez_compressed(os.path.join(OUT, "synthetic_big.npz"), **b)
        X, y = synth(4096, 32)
        m = pyGPs.GPR()
        m.setPrior(kernel=pyGPs.cov.RBFard(D=32, log_ell_list=[np.log(3.0)] * 32, log_sigma=0.0))
        nlZ, post = m.getPosterior(X, y, der=False)
        b["c3_4096_nlZ"] = np.float64(nlZ)
        b["c3_4096_alpha"] = post.alpha
        print("c3 big", repr(nlZ), flush=True)
        np.savez_compressed(os.path.join(OUT, "synthetic_big.npz"), **b)


def housing():
    """The one published number: doc/source/demoHousing.rst:30 (optimized nlZ 214.46);
    preprocessing per Demo/Housing/demo_Housing.py:25-40."""
    path = os.path.join(REF, "pyGPs/Demo/Housing/housing.txt")
    data = np.genfromtxt(path)
    N = 25
    x = np.concatenate((data[:-N, :4], data[:-N, 5:-1]), axis=1)
    x = (x - np.mean(x, axis=0)) / (np.std(x, axis=0) + 1.e-16)
    y = np.reshape(data[:-N, -1], (len(data[:-N, -1]), 1))
    y = (y - np.mean(y)) / (np.std(y) + 1.e-16)
    x_train, y_train = x, y
    s = {"x": x_train, "y": y_train}
    m = pyGPs.GPR()
    nlZ, dn, post = m.getPosterior(x_train, y_train)
    s["default_nlZ"] = np.float64(nlZ)
    s["default_dcov"] = np.array(dn.cov)
    s["default_dlik"] = np.array(dn.lik)
    m = pyGPs.GPR()
    m.optimize(x_train, y_train)
    s["opt_nlZ"] = np.float64(m.nlZ)
    s["opt_hyp"] = np.array(m.covfunc.hyp + m.likfunc.hyp)
    print("housing", s["default_nlZ"], s["opt_nlZ"], s["opt_hyp"])
    np.savez_compressed(os.path.join(OUT, "housing.npz"), **s)




def classification():
    """Config 5 family: GPC + inf.EP + lik.Erf.  KAT5 on the reference's Demo/Classification fixture
    (nlZ=50.454379530956) and synthetic sign labels at N=200/512, D=16, RBF(log 4, 0)."""
    d = np.load(os.path.join(REF, "pyGPs/Demo/Classification/classification_data.npz"))
    x, y, xs = d["x"], d["y"], d["xstar"][::40]
    s = {"kat5_x": x, "kat5_y": y, "kat5_xs": xs}
    m = pyGPs.GPC()
    nlZ, dn, post = m.getPosterior(x, y)
    s["kat5_nlZ"] = np.float64(nlZ); s["kat5_dcov"] = np.array(dn.cov); s["kat5_alpha"] = post.alpha
    s["kat5_sW"] = post.sW; s["kat5_L"] = post.L
    s["kat5_ttau"] = m.inffunc.last_ttau; s["kat5_tnu"] = m.inffunc.last_tnu
    out = m.predict(xs, np.ones((xs.shape[0], 1)))
    for name, v in zip(("ym", "ys2", "fm", "fs2", "lp"), out):
        s["kat5_" + name] = v
    print("kat5 nlZ", nlZ, dn.cov)
    for N in (200, 512):
        rng = np.random.default_rng(0)
        X = rng.standard_normal((N, 16))
        lab = np.sign(X[:, :1] + 0.5 * X[:, 1:2] + 0.3 * rng.standard_normal((N, 1)))
        lab[lab == 0] = 1
        m = pyGPs.GPC()
        m.setPrior(kernel=pyGPs.cov.RBF(np.log(4.0), 0.0))
        nlZ, dn, post = m.getPosterior(X, lab)
        Xs = np.random.default_rng(1).standard_normal((64, 16))
        out = m.predict(Xs)
        tag = "c5_%d" % N
        s[tag + "_nlZ"] = np.float64(nlZ); s[tag + "_dcov"] = np.array(dn.cov); s[tag + "_alpha"] = post.alpha
        s[tag + "_sW"] = post.sW; s[tag + "_ym"] = out[0]; s[tag + "_ys2"] = out[1]; s[tag + "_fm"] = out[2]
        s[tag + "_fs2"] = out[3]; s[tag + "_lp"] = out[4] if out[4] is not None else np.zeros(0)
        print(tag, nlZ)
    np.savez_compressed(os.path.join(OUT, "classification.npz"), **s)


def cov_programs():
    """The kernels evaluated by the device program (csrc/covprog.cu) and composites of them: matrices in all three
    modes and every derivative matrix from the reference's own classes, plus the Mauna Loa model of
    Demo/MaunaLoa/demo_MaunaLoa.py:36-68 (data loaded exactly as the demo does)."""
    c = pyGPs.cov
    rng = np.random.RandomState(3)
    x3 = rng.normal(0, 1.0, (25, 3)); z3 = rng.normal(0, 1.0, (11, 3)); z3[4] = x3[7]     # one coinciding point (Noise)
    x1 = rng.normal(0, 2.0, (30, 1)); z1 = rng.normal(0, 2.0, (9, 1))
    s = {"x3": x3, "z3": z3, "x1": x1, "z1": z1}
    leaves = {
        "rbfunit": (lambda: c.RBFunit(0.3), 3),
        "rq": (lambda: c.RQ(0.2, -0.1, 0.4), 3),
        "rqard": (lambda: c.RQard(log_ell_list=[0.1, -0.2, 0.3], log_sigma=0.2, log_alpha=-0.3), 3),
        "periodic": (lambda: c.Periodic(0.2, 0.5, 0.1), 1),
        "piecepoly0": (lambda: c.PiecePoly(0.9, 0, 0.1), 3),
        "piecepoly1": (lambda: c.PiecePoly(0.9, 1, 0.1), 3),
        "piecepoly2": (lambda: c.PiecePoly(0.9, 2, 0.1), 3),
        "piecepoly3": (lambda: c.PiecePoly(1.2, 3, -0.2), 3),
        "gabor": (lambda: c.Gabor(0.4, 0.3), 3),
        "noise": (lambda: c.Noise(-0.7), 3),
        "const": (lambda: c.Const(0.3), 3),
        "linear": (lambda: c.Linear(-0.4), 3),
        "poly": (lambda: c.Poly(0.2, 3, -0.3), 3),
        "comp3": (lambda: c.RBF(0.3, 0.1) * c.Matern(0.5, 5, -0.2) + c.RQ(0.2, -0.1, 0.4) * 1.5 + c.Noise(-1.0)
                  + c.Const(-0.5) * c.Linear(-1.0), 3),
        "comp1": (lambda: c.RBF(1.0, 0.5) + c.Periodic(0.2, 0.5, 0.1) * c.RBF(1.5, -0.2) + c.RQ(0.1, -0.4, -0.2)
                  + (c.RBF(-1.0, -1.0) + c.Noise(-1.5)), 1),
    }
    for name, (make, D) in leaves.items():
        k = make()
        x, z = (x3, z3) if D == 3 else (x1, z1)
        s[name + "_hyp"] = np.array(k.hyp, dtype=float)
        s[name + "_train"] = k.getCovMatrix(x=x, mode="train")
        s[name + "_cross"] = k.getCovMatrix(x=x, z=z, mode="cross")
        s[name + "_self"] = k.getCovMatrix(z=z, mode="self_test")
        for i in range(len(k.hyp)):
            s["%s_dtrain%d" % (name, i)] = k.getDerMatrix(x=x, mode="train", der=i)
            s["%s_dcross%d" % (name, i)] = k.getDerMatrix(x=x, z=z, mode="cross", der=i)
            s["%s_dself%d" % (name, i)] = k.getDerMatrix(z=z, mode="self_test", der=i)
    # Mauna Loa (Demo/MaunaLoa/demo_MaunaLoa.py:36-68)
    year, co2 = [], []
    for line in open(os.path.join(REF, "pyGPs/Demo/MaunaLoa/mauna.txt")):
        zz = line.split('  ')
        v = float(zz[1].split('\n')[0])
        if v != -99.99:
            year.append(float(zz[0])); co2.append(v)
    X = np.array([i for i, j in zip(year, co2) if i < 2004]).reshape(-1, 1)
    Y = np.array([j for i, j in zip(year, co2) if i < 2004]).reshape(-1, 1)
    xs = np.arange(2004 + 1. / 24., 2024 - 1. / 24., 1. / 12.).reshape(-1, 1)
    k1 = c.RBF(np.log(67.), np.log(66.))
    k2 = c.Periodic(np.log(1.3), np.log(1.0), np.log(2.4)) * c.RBF(np.log(90.), np.log(2.4))
    k3 = c.RQ(np.log(1.2), np.log(0.66), np.log(0.78))
    k4 = c.RBF(np.log(1.6 / 12.), np.log(0.18)) + c.Noise(np.log(0.19))
    k = k1 + k2 + k3 + k4
    m = pyGPs.GPR()
    m.setData(X, Y)
    m.setPrior(kernel=k)
    s["mauna_x"], s["mauna_y"], s["mauna_xs"] = X, Y, xs
    s["mauna_c"] = np.float64(m.meanfunc.hyp[0])
    run_model(m, X, Y, xs, "mauna", s, keep_L=False)
    print("mauna nlZ", repr(s["mauna_nlZ"]), s["mauna_dcov"])
    # cov.Pre: the same matrices handed over precomputed (GraphExtensions-style use, Core/cov.py:1429-1455)
    kk = c.RBF(0.3, 0.1)
    M2 = kk.getCovMatrix(x=x3, mode="train")
    M1 = np.vstack([kk.getCovMatrix(x=x3, z=z3, mode="cross"), kk.getCovMatrix(z=z3, mode="self_test").T])
    y3 = np.sin(x3.sum(1, keepdims=True))
    m = pyGPs.GPR()
    m.setPrior(kernel=c.Pre(M1, M2))
    nlZ, post = m.getPosterior(x3, y3, der=False)
    s["pre_M1"], s["pre_M2"], s["pre_y"] = M1, M2, y3
    s["pre_nlZ"], s["pre_alpha"] = np.float64(nlZ), post.alpha
    out = m.predict(z3)
    s["pre_ym"], s["pre_ys2"] = out[0], out[1]
    m = pyGPs.GPR()
    m.setPrior(kernel=c.Pre(M1, M2) + c.Noise(-1.0))
    nlZ, dn, post = m.getPosterior(x3, y3)
    s["prenoise_nlZ"], s["prenoise_dcov"], s["prenoise_dlik"] = np.float64(nlZ), np.array(dn.cov), np.array(dn.lik)
    s["prenoise_alpha"] = post.alpha
    np.savez_compressed(os.path.join(OUT, "cov_programs.npz"), **s)


def c4_big():
    """SURVEY 8(c)/(d)'s larger pins of config 4: (N,M) = (32768,512) -> nlZ 85900.043566, (65536,1024) -> 170971.078457."""
    s = {}
    for N, M in ((32768, 512), (65536, 1024)):
        rng = np.random.default_rng(0)
        X = rng.standard_normal((N, 8))
        y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
        U = rng.standard_normal((M, 8))
        Xs = np.random.default_rng(1).standard_normal((300, 8))
        m = pyGPs.GPR_FITC()
        m.setPrior(kernel=pyGPs.cov.RBF(np.log(2.0), 0.0), inducing_points=U)
        run_model(m, X, y, Xs, "c4_%d_%d" % (N, M), s, keep_L=False)
        print("c4 big", N, M, repr(s["c4_%d_%d_nlZ" % (N, M)]), flush=True)
        np.savez_compressed(os.path.join(OUT, "synthetic_c4big.npz"), **s)


def c5_big():
    """SURVEY 8(c)/(d)'s pin of config 5: N=1024, D=16, RBF(log 4, 0), Erf -> nlZ 344.43995473 in 4 sweeps."""
    s = {}
    N = 1024
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, 16))
    lab = np.sign(X[:, :1] + 0.5 * X[:, 1:2] + 0.3 * rng.standard_normal((N, 1)))
    lab[lab == 0] = 1
    m = pyGPs.GPC()
    m.setPrior(kernel=pyGPs.cov.RBF(np.log(4.0), 0.0))
    nlZ, dn, post = m.getPosterior(X, lab)
    Xs = np.random.default_rng(1).standard_normal((64, 16))
    out = m.predict(Xs)
    tag = "c5_%d" % N
    s[tag + "_nlZ"] = np.float64(nlZ); s[tag + "_dcov"] = np.array(dn.cov); s[tag + "_alpha"] = post.alpha
    s[tag + "_sW"] = post.sW; s[tag + "_ym"] = out[0]; s[tag + "_ys2"] = out[1]; s[tag + "_fm"] = out[2]
    s[tag + "_fs2"] = out[3]
    print(tag, repr(nlZ), flush=True)
    np.savez_compressed(os.path.join(OUT, "classification_c5big.npz"), **s)


if __name__ == "__main__":
    if "--programs" in sys.argv:
        cov_programs()
    elif "--c4big" in sys.argv or "--c5big" in sys.argv:
        if "--c5big" in sys.argv:
            c5_big()
        if "--c4big" in sys.argv:
            c4_big()
    elif "--only-classification" in sys.argv:
        classification()
    else:
        kat_regression()
        cov_vectors()
        synthetic(big="--big" in sys.argv)
        housing()
        classification()
