"""Host-side logic (optimizers, containers, plugin classes) checked on the CPU by driving it
with the oracle as the objective - no GPU, no libgpk compute calls."""
import copy

import numpy as np
import pytest

import pygps_b200 as pg
from pygps_b200 import opt
from oracle import gp_oracle as go


class _OracleModel(object):
    """A stand-in model whose getPosterior runs the CPU oracle; exercises Optimizer plumbing."""

    def __init__(self, x, y, c):
        self.x, self.y = x, y
        self.meanfunc = pg.mean.Const(c)
        self.covfunc = pg.cov.RBF()
        self.likfunc = pg.lik.Gauss()
        self.calls = 0

    def getPosterior(self, der=True):
        self.calls += 1
        mean = ("const", self.meanfunc.hyp[0])
        covs = ("rbf", list(self.covfunc.hyp))
        if not der:
            post, nlZ = go.exact_evaluate(mean, covs, self.likfunc.hyp[0], self.x, self.y, 2)
            return nlZ, post
        post, nlZ, dn = go.exact_evaluate(mean, covs, self.likfunc.hyp[0], self.x, self.y, 3)
        d = pg.inf.dnlZStruct(self.meanfunc, self.covfunc, self.likfunc)
        d.mean, d.cov, d.lik = list(dn["mean"]), list(dn["cov"]), list(dn["lik"])
        return nlZ, d, post


def test_minimize_reproduces_the_reference_optimum(golden):
    """KAT2: pyGPs.GPR().setData(x,y); optimize(x,y) -> nlZ 11.5714075328."""
    g = golden("kat_regression")
    m = _OracleModel(g["x"], g["y"], float(g["kat2_c"]))
    o = opt.Minimize(m)
    hyp, val = o.findMin(m.x, m.y, numIters=40)
    assert abs(val - float(g["kat2_opt_nlZ"])) < 1e-6 * abs(float(g["kat2_opt_nlZ"]))
    np.testing.assert_allclose(hyp, g["kat2_opt_hyp"], rtol=1e-5, atol=1e-6)
    assert abs(val - 11.5714075328) < 1e-6


def test_minimize_on_rosenbrock_and_eval_budget():
    def f(v):
        x, y = v
        return (1 - x) ** 2 + 100 * (y - x * x) ** 2, np.array([-2 * (1 - x) - 400 * x * (y - x * x), 200 * (y - x * x)])
    X, fx, it = opt.minimize(f, np.array([-1.2, 1.0]), length=200)
    assert fx[-1] < 1e-10 and np.allclose(X, [1, 1], atol=1e-4)
    assert all(b <= a + 1e-12 for a, b in zip(fx, fx[1:]))
    n = [0]

    def g(v):
        n[0] += 1
        return f(v)
    opt.minimize(g, np.array([-1.2, 1.0]), length=-25)
    assert n[0] <= 26


def test_scg_descends(golden):
    g = golden("kat_regression")
    m = _OracleModel(g["x"], g["y"], float(g["kat2_c"]))
    before = m.getPosterior(der=False)[0]
    hyp, val = opt.SCG(m).findMin(m.x, m.y, numIters=30)
    assert val < before


@pytest.mark.parametrize("cls", [opt.CG, opt.BFGS, opt.Simplex])
def test_scipy_backed_optimizers_descend(golden, cls):
    """The reference's only semantic test: funcValue < nlZ before optimisation (Testing/unit_test_opt.py:39)."""
    g = golden("kat_regression")
    m = _OracleModel(g["x"], g["y"], float(g["kat2_c"]))
    before = m.getPosterior(der=False)[0]
    hyp, val = cls(m).findMin(m.x, m.y, numIters=15)
    assert val < before and len(hyp) == 4


def test_random_restarts_stop_conditions(golden):
    g = golden("kat_regression")
    m = _OracleModel(g["x"], g["y"], float(g["kat2_c"]))
    conf = opt.random_init_conf(m.meanfunc, m.covfunc, m.likfunc)
    conf.num_restarts = 3
    conf.covRange = [(-1, 1), (-1, 1)]
    with pytest.raises(Exception):
        conf.covRange = [(-1, 1)]
    np.random.seed(0)
    o = opt.Minimize(m, conf)
    hyp, val = o.findMin(m.x, m.y, numIters=10)
    assert o.trailsCounter == 3 and val < 60.0


def test_hyp_flattening_order():
    m = pg.GPR()
    m.setPrior(mean=pg.mean.Linear(D=2), kernel=pg.cov.RBFard(D=2))
    o = m.optimizer
    o.model = m
    arr = o._convert_to_array()
    assert arr.tolist() == [0.5, 0.5, 0.0, 0.0, 0.0, np.log(0.1)]
    o._apply_in_objects(np.arange(6.0))
    assert m.meanfunc.hyp == [0.0, 1.0] and m.covfunc.hyp == [2.0, 3.0, 4.0] and m.likfunc.hyp == [5.0]


def test_kernel_operator_overloads_and_hyp_plumbing():
    k = pg.cov.RBF(0.1, 0.2) + pg.cov.Matern(0.3, 5, 0.4)
    assert isinstance(k, pg.cov.SumOfKernel) and k.hyp == [0.1, 0.2, 0.3, 0.4]
    k.hyp = [1., 2., 3., 4.]
    assert k.cov1.hyp == [1., 2.] and k.cov2.hyp == [3., 4.]
    p = pg.cov.RBF() * pg.cov.RBFard(D=2)
    assert isinstance(p, pg.cov.ProductOfKernel) and len(p.hyp) == 5
    s = pg.cov.RBF() * 3.0
    assert isinstance(s, pg.cov.ScaleOfKernel) and s.hyp == [3.0, 0., 0.]
    f = pg.cov.RBF().fitc(np.zeros((3, 2)))
    assert isinstance(f, pg.cov.FITCOfKernel)
    f.hyp = [0.5, 0.6]
    assert f.covfunc.hyp == [0.5, 0.6]
    with pytest.raises(Exception):
        pg.cov.RBF().getCovMatrix(x=np.zeros((2, 1)))            # mode missing
    with pytest.raises(Exception):
        pg.cov.RBF().getCovMatrix(x=np.zeros((2, 1)), mode='cross')
    with pytest.raises(Exception):
        pg.cov.RBF().getDerMatrix(x=np.zeros((2, 1)), mode='train')
    assert pg.cov.Matern(d=4)._matern_d() == 3 and pg.cov.Matern(d=7.0)._matern_d() == 7


def test_means_and_gauss_likelihood_host_math():
    x = np.arange(6.0).reshape(3, 2)
    assert pg.mean.Zero().getMean(x).shape == (3, 1)
    np.testing.assert_allclose(pg.mean.Const(2.).getMean(x), 2 * np.ones((3, 1)))
    np.testing.assert_allclose(pg.mean.Linear(D=2).getMean(x), x.sum(1, keepdims=True) * 0.5)
    s = pg.mean.Const(2.) + pg.mean.Linear(D=2)
    assert s.hyp == [2., 0.5, 0.5]
    np.testing.assert_allclose(s.getDerMatrix(x, 2), x[:, 1:2])
    sc = pg.mean.Linear(D=2) * 3.0
    np.testing.assert_allclose(sc.getMean(x), 1.5 * x.sum(1, keepdims=True))
    lk = pg.lik.Gauss(np.log(0.3))
    mu, s2, y = np.array([[0.1], [0.2]]), np.array([[0.5], [0.0]]), np.array([[0.0], [1.0]])
    lp, ymu, ys2 = lk.evaluate(y, mu, s2, None, None, 3)
    np.testing.assert_allclose(ys2, s2 + 0.09)
    np.testing.assert_allclose(lp, -(y - mu) ** 2 / (0.09 + s2) / 2 - np.log(2 * np.pi * (0.09 + s2)) / 2)


def test_model_defaults_match_the_reference():
    m = pg.GPR()
    assert isinstance(m.meanfunc, pg.mean.Zero) and isinstance(m.covfunc, pg.cov.RBF)
    assert isinstance(m.inffunc, pg.inf.Exact) and isinstance(m.optimizer, pg.opt.Minimize)
    assert m.likfunc.hyp == [np.log(0.1)]
    x = np.linspace(0, 1, 7)
    y = np.sin(x)
    m.setData(x, y)                                       # 1-d inputs become columns, Zero -> Const(mean(y))
    assert m.x.shape == (7, 1) and m.y.shape == (7, 1)
    assert isinstance(m.meanfunc, pg.mean.Const) and abs(m.meanfunc.hyp[0] - y.mean()) < 1e-15
    f = pg.GPR_FITC()
    with pytest.raises(Exception):
        f.setPrior(kernel=pg.cov.RBF())                   # no inducing points yet
    f.setData(np.random.rand(10, 2), np.random.rand(10))
    assert f.u.shape == (25, 2) and isinstance(f.covfunc, pg.cov.FITCOfKernel)
    with pytest.raises(AssertionError):
        pg.GPR().setData(np.zeros((3, 1)), np.zeros((4, 1)))
    with pytest.raises(Exception):
        pg.inf.Exact().evaluate(pg.mean.Zero(), pg.cov.RBF(), object(), np.zeros((2, 1)), np.zeros((2, 1)))


def test_poststruct_deepcopy_keeps_lazy_factor_and_dnlz_accumulates():
    p = pg.inf.postStruct()
    p.alpha = np.ones((3, 1)); p.sW = np.ones((3, 1)); p.L = np.eye(3)
    q = copy.deepcopy(p)
    assert q.L is not p.L and np.array_equal(q.L, p.L)
    m, c, l = pg.mean.Const(1.), pg.cov.RBF(), pg.lik.Gauss()
    d1, d2 = pg.inf.dnlZStruct(m, c, l), pg.inf.dnlZStruct(m, c, l)
    d1.cov, d2.cov = [1., 2.], [3., 4.]
    assert d1.accumulateDnlZ(d2).cov == [4., 6.] and len(d1.mean) == 1 and len(d1.lik) == 1


def test_bench_clock_sampler_windows_and_reference_arm_line():
    """bench.py host logic: the clock summary is taken from the samples of the timed region (falling back, flagged, to the
    warm-up's when the region holds none), throttle reasons are collected, and `--impl reference` prints ONE JSON line
    with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench

    class FakeProc(object):
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    def line(mhz, cap="Not Active"):
        return "0, %d, 1965, 400.0, 0x0, Not Active, Not Active, Not Active, %s" % (mhz, cap)
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.lines = [line(1200), line(1900), line(1950, "Active"), line(1960), line(1300)]
    out = s.stop(1, 4)
    assert out["samples"] == 3 and out["sm_mhz"] == 1950.0 and out["reasons"] == ["sw_power_cap"] and "note" not in out
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.lines = [line(1900), line(1910)]
    out = s.stop(2, 2)                                   # empty timed window: warm-up samples, flagged
    assert out["samples"] == 2 and "note" in out
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--problem-n", "512"], capture_output=True, text=True, timeout=300)
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert p.returncode == 0 and len(lines) == 1, (p.returncode, p.stdout[-300:], p.stderr[-300:])
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    # the reference arm reports what it actually ran (a measured full-size evaluation, never an extrapolation) and
    # the SAME config object as the GPU arm
    assert d["steps"] >= 1 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6
    assert d["config"] == bench.make_config(512, 8, 1)


def test_random_restarts_run_concurrently_with_serial_bookkeeping(monkeypatch):
    """Core/opt.py:301-327: the random restarts are independent minimisations.  With several devices they run on one
    model copy per device, concurrently; starts are drawn and results accounted for in the serial order, so the outcome
    (best hyper-parameters, value, trial / error counters) is identical to a one-device run."""
    import threading
    import pygps_b200 as pg
    from pygps_b200 import opt

    seen = {}

    class Quad(opt.Optimizer):
        _label, _failtext = 'Quad', 'quad'

        def findMin(self, x, y, numIters=10):
            return self._search(self._convert_to_array(), numIters)

        def _run_once(self, hyp, numIters, first):
            seen.setdefault(threading.get_ident(), []).append(tuple(self.model.devices or [None]))
            import time
            time.sleep(0.05)                     # an evaluation takes a while: the other devices pick up trials
            h = np.array(hyp, dtype=float)
            if abs(h[1]) > 4.0:
                raise RuntimeError("trial failed")
            return h * 0.5, float(np.sum((h - 1.0) ** 2))

    def run(devs):
        np.random.seed(7)
        m = pg.GPR()
        m.setDevices(devs)
        conf = opt.random_init_conf(m.meanfunc, m.covfunc, m.likfunc)
        conf.num_restarts = 9
        o = Quad(m, conf)
        m.optimizer = o
        monkeypatch.setattr(Quad, "_restart_devices", lambda self: list(devs))
        res = o.findMin(None, None)
        return res, o.trailsCounter, o.errorCounter, getattr(o, "devices_used", None)

    seen.clear()
    (h1, v1), t1, e1, used1 = run([0])
    n_threads_serial = len(seen)
    seen.clear()
    (h3, v3), t3, e3, used3 = run([0, 1, 2])
    assert np.array_equal(h1, h3) and v1 == v3 and (t1, e1) == (t3, e3) and t1 == 9
    assert n_threads_serial == 1 and len(seen) == 3 and used3 == [0, 1, 2]
    assert sorted(set(d for v in seen.values() for d in v)) == [(0,), (1,), (2,)]
    # min_threshold only: waves of one trial per device until the threshold is met
    np.random.seed(3)
    m = pg.GPR()
    conf = opt.random_init_conf(m.meanfunc, m.covfunc, m.likfunc)
    conf.min_threshold = 4.0
    o = Quad(m, conf)
    m.optimizer = o
    monkeypatch.setattr(Quad, "_restart_devices", lambda self: [0, 1])
    h, v = o.findMin(None, None)
    assert v <= 4.0


def test_models_with_engines_survive_deepcopy_and_pickle():
    """reference models are plain Python objects; ours must stay copyable / picklable after evaluation."""
    import copy
    import pickle
    import pygps_b200 as pg
    from pygps_b200 import _lib

    class FakeEngine(_lib.Engine):
        def __init__(self):                     # no GPU here: only the copy protocol is under test
            self.epoch = 3
    m = pg.GPR()
    m.inffunc._engine = FakeEngine()
    p = pg.inf.postStruct()
    p.alpha, p.sW, p.L = np.ones((3, 1)), np.ones((3, 1)), np.eye(3)
    m.posterior = p
    m2 = copy.deepcopy(m)
    assert m2.inffunc._engine is None and np.array_equal(m2.posterior.L, np.eye(3))
    m3 = pickle.loads(pickle.dumps(m))
    assert m3.inffunc._engine is None and np.array_equal(m3.posterior.L, np.eye(3))


def test_covariance_programs_are_emitted_in_post_order_with_the_reference_hyp_layout():
    """cov.Kernel._device_prog(): nodes in post-order (children before parents, root last), every node's hyp0 the index
    of its first own entry in the composite's flat hyper-parameter list exactly as the reference concatenates it
    (Core/cov.py:235, 270, 303-306: cov1.hyp + cov2.hyp; ScaleOfKernel: [scalar] + cov.hyp)."""
    import pygps_b200 as pg
    from pygps_b200 import _lib
    c = pg.cov
    k = c.RBF(0.1, 0.2) * c.Periodic(0.3, 0.4, 0.5) + c.Noise(0.6)
    nodes, hyp = k._device_prog()
    assert hyp == [0.1, 0.2, 0.3, 0.4, 0.5, 0.6] == k.hyp
    assert nodes == [(_lib.OP_RBF, -1, -1, 0, 0.0), (_lib.OP_PERIODIC, -1, -1, 2, 0.0), (_lib.OP_PROD, 0, 1, -1, 0.0),
                     (_lib.OP_NOISE, -1, -1, 5, 0.0), (_lib.OP_SUM, 2, 3, -1, 0.0)]
    # scalar * kernel: the scalar is the node's own hyper-parameter, the child's follow it
    k = 3.0 * c.Matern(0.7, 5, 0.8) + c.RQard(log_ell_list=[0.1, 0.2], log_sigma=0.3, log_alpha=0.4)
    nodes, hyp = k._device_prog()
    assert hyp == [3.0, 0.7, 0.8, 0.1, 0.2, 0.3, 0.4]
    assert nodes[0] == (_lib.OP_MATERN, -1, -1, 1, 5.0) and nodes[1] == (_lib.OP_SCALE, 0, -1, 0, 0.0)
    assert nodes[2] == (_lib.OP_RQARD, -1, -1, 3, 0.0) and nodes[3] == (_lib.OP_SUM, 1, 2, -1, 0.0)
    # setting the composite's hyp propagates into the parts (what the optimizers do every iteration)
    k.hyp = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0]
    assert k.cov1.cov.hyp == [2.0, 3.0] and k.cov2.hyp == [4.0, 5.0, 6.0, 7.0] and k._device_prog()[1][0] == 1.0
    # PiecePoly / Poly carry their integer parameter; precomputed matrices are reported as Pre leaves
    assert c.PiecePoly(0.1, 3, 0.2)._device_prog()[0][0][4] == 3.0 and c.Poly(0.1, 4, 0.2)._device_prog()[0][0][4] == 4.0
    pre = c.Pre(np.ones((4, 2)), np.eye(3))
    kk = pre + c.RBF()
    assert kk._pre_leaves() == [pre] and kk._device_prog()[0][0][0] == _lib.OP_PRE
    assert np.array_equal(pre.getCovMatrix(mode='train'), np.eye(3)) and pre.getCovMatrix(mode='cross').shape == (3, 2)
    assert pre.getCovMatrix(mode='self_test').shape == (2, 1)

    class Mine(c.Kernel):                       # a kernel without a device implementation: no program, no CPU fallback
        def __init__(self):
            self.hyp = [0.0]
    assert (Mine() + c.RBF())._device_prog() is None
    # the native single kernels keep their dedicated fused build
    assert c.RBF()._device_spec() is not None and (c.RBF() + c.RBF())._device_spec() is None
