"""Multi-GPU through the plugin API: random restarts spread over the visible GPUs, and ONE
evaluation sharded over them in-process (one thread + one libgpk handle per GPU, NCCL underneath).  On a one-GPU box
the same code paths run with world size 1 (no NCCL): gpk_exact_eval_dist, gpk_dist_gather_factor,
gpk_exact_eval_dist_der and the sharded predict are all exercised; with 2+ GPUs the exchanges are real."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import pygps_b200 as pg                    # noqa: E402
from pygps_b200 import _lib                # noqa: E402
from parity_report import check            # noqa: E402


def _data(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    return X, y


def test_optimize_spreads_random_restarts_over_all_visible_gpus():
    devs = _lib.visible_devices()
    X, y = _data(400, 3)
    results = {}
    for name, use in (("all", devs), ("one", devs[:1])):
        np.random.seed(11)
        m = pg.GPR()
        m.setDevices(use)
        m.setOptimizer("Minimize", num_restarts=max(4, 2 * len(devs)), covRange=[(-1, 1), (-1, 1)], likRange=[(-3, 0)])
        m.optimize(X, y, numIterations=15)
        results[name] = (m.nlZ, list(m.covfunc.hyp) + list(m.likfunc.hyp), m.optimizer.devices_used,
                         m.optimizer.trailsCounter)
    assert results["all"][2] == sorted(devs), "restarts did not reach every visible GPU: %r" % (results["all"][2],)
    assert results["one"][2] == devs[:1]
    assert results["all"][3] == results["one"][3] == max(4, 2 * len(devs))
    check("parallel vs serial restarts: nlZ", results["all"][0], results["one"][0], 1e-9)
    check("parallel vs serial restarts: hyp", results["all"][1], results["one"][1], 1e-6)


@pytest.mark.parametrize("N,D,kern", [(700, 4, "rbf"), (4500, 6, "ard"), (3000, 5, "matern")])
def test_sharded_model_matches_single_gpu_model(N, D, kern):
    """GPR(shard=True): getPosterior with and without derivatives, post.L and predict through the sharded engine against
    the single-GPU engine (N=4500: int8 blocked variant of the sharded factorisation, ragged last panel)."""
    devs = _lib.visible_devices()
    X, y = _data(N, D, seed=N)
    Xs = np.random.default_rng(5).standard_normal((333, D))

    def kernel():
        if kern == "rbf":
            return pg.cov.RBF(np.log(1.7), 0.1)
        if kern == "ard":
            return pg.cov.RBFard(log_ell_list=list(np.linspace(0.2, 0.9, D)), log_sigma=0.2)
        return pg.cov.Matern(np.log(1.5), 5, 0.1)
    ref = pg.GPR(devices=devs[:1], shard=False)
    ref.setPrior(kernel=kernel())
    sh = pg.GPR(devices=devs, shard=True)
    sh.setPrior(kernel=kernel())
    n0, d0, p0 = ref.getPosterior(X, y)
    n1, d1, p1 = sh.getPosterior(X, y)
    assert isinstance(sh.inffunc._sharded, _lib.ShardedEngine) and sh.inffunc._sharded.world == len(devs)
    tag = "sharded x%d %s N=%d " % (len(devs), kern, N)
    check(tag + "nlZ", n1, n0, 1e-10)
    check(tag + "alpha", p1.alpha, p0.alpha, 1e-8)
    check(tag + "dnlZ.cov", d1.cov, d0.cov, 1e-8)
    check(tag + "dnlZ.lik", d1.lik, d0.lik, 1e-8)
    out1 = sh.predict(Xs)
    out0 = ref.predict(Xs)
    check(tag + "ym", out1[0], out0[0], 1e-8)
    check(tag + "ys2", out1[1], out0[1], 1e-8)
    if N <= 1000:
        check(tag + "post.L", p1.L, p0.L, 1e-9)
    n2, p2 = sh.getPosterior(X, y, der=False)
    check(tag + "nlZ (der=False)", n2, n0, 1e-10)


def test_evaluation_is_routed_to_the_sharded_path_when_the_factor_does_not_fit(monkeypatch):
    """inf.Exact routes by memory: pretend the GPU is tiny and check that a sharded engine is used (and right)."""
    devs = _lib.visible_devices()
    X, y = _data(600, 3)
    m = pg.GPR(devices=devs + devs[:1] if len(devs) == 1 else devs)      # the routing needs a list of >= 2 entries
    monkeypatch.setattr(_lib, "device_memory", lambda d: (1 << 20, 1 << 20))
    assert m.inffunc._wants_sharding(600)
    monkeypatch.setattr(_lib, "device_memory", lambda d: (150 << 30, 180 << 30))
    assert not m.inffunc._wants_sharding(16384)
    assert m.inffunc._wants_sharding(150000)
    monkeypatch.undo()
    free, total = _lib.device_memory(devs[0])
    assert 0 < free <= total and total > (64 << 30)
