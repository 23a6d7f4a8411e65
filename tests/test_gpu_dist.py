"""The evaluation sharded by block columns over several GPUs (gpk_exact_eval_dist), one process per GPU.
World size 1 runs on any GPU box (no NCCL needed); world size 2 needs two GPUs and is skipped otherwise."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, math, os, sys
    sys.path.insert(0, %r)
    import numpy as np
    from pygps_b200 import _lib
    from pygps_b200._dist import DistCtx
    ctx = DistCtx()
    eng = _lib.Engine(ctx.local_rank)
    ctx.shard_engine(eng)
    out = {}
    for N, D in ((300, 3), (2048, 8), (4096, 8), (4500, 5)):
        rng = np.random.default_rng(0)
        X = rng.standard_normal((N, D))
        y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
        eng.set_data(X)
        nlZ, alpha = eng.exact_eval_dist(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), y.reshape(-1))
        out[str(N)] = [float(nlZ), float(np.abs(alpha).sum()), float(alpha[7, 0])]
    ctx.barrier()
    print("RESULT", ctx.rank, json.dumps(out))
    ctx.close()
""") % ROOT


def _gpu_count():
    import ctypes
    from pygps_b200 import _lib
    c = ctypes.c_int(0)
    _lib.load().gpk_device_count(ctypes.byref(c))
    return c.value


def _reference(N, D):
    from pygps_b200 import _lib
    import math
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    e = _lib.Engine(0)
    e.set_data(X)
    nlZ, alpha, _, _ = e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), y.reshape(-1), False)
    return float(nlZ), float(np.abs(alpha).sum()), float(alpha[7, 0])


def _run(world, tmp_path, port, extra_env=None):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    res = {}
    for o in outs:
        for line in o[0].splitlines():
            if line.startswith("RESULT"):
                _, rank, payload = line.split(" ", 2)
                res[int(rank)] = json.loads(payload)
    return res


def _check_sharded(world, blocked, tmp_path, golden):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    res = _run(world, tmp_path, 29560 + world + 10 * blocked, {"GPK_DIST_OZAKI": str(blocked)})
    assert len(res) == world
    g = golden("synthetic")
    for N, D in ((300, 3), (2048, 8), (4096, 8), (4500, 5)):
        ref = _reference(N, D)
        for rank in range(world):
            got = res[rank][str(N)]
            assert abs(got[0] - ref[0]) < 1e-10 * abs(ref[0]), (world, rank, N, got, ref)
            assert abs(got[1] - ref[1]) < 1e-8 * abs(ref[1]) and abs(got[2] - ref[2]) < 1e-7 * max(1, abs(ref[2]))
    assert abs(res[0]["2048"][0] - float(g["c2_2048_nlZ"])) < 1e-9 * abs(float(g["c2_2048_nlZ"]))


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_eval_matches_single_gpu(world, tmp_path, golden):
    _check_sharded(world, 0, tmp_path, golden)


FITC_WORKER = textwrap.dedent("""
    import json, math, os, sys
    sys.path.insert(0, %r)
    import numpy as np
    from pygps_b200 import _lib
    from pygps_b200._dist import DistCtx
    ctx = DistCtx()
    eng = _lib.Engine(ctx.local_rank)
    ctx.shard_engine(eng)
    N, M, D = 3000, 150, 4
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, D)); y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    U = rng.standard_normal((M, D))
    lo, hi = (N * ctx.rank) // ctx.world, (N * (ctx.rank + 1)) // ctx.world      # this rank's rows
    eng.set_data(X[lo:hi])
    nlZ, alpha, Lp, dcov, dlik, al = eng.fitc_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), U,
                                                   y[lo:hi].reshape(-1), True)
    Xs = np.random.default_rng(1).standard_normal((40, D))
    ka, fs2 = eng.fitc_predict(Xs)
    ctx.barrier()
    print("RESULT", ctx.rank, json.dumps({"nlZ": float(nlZ), "alpha": alpha[:, 0].tolist(), "dcov": list(map(float, dcov)),
          "dlik": float(dlik[0]), "Ltrace": float(np.trace(Lp)), "ka": ka[:, 0].tolist(), "fs2": fs2[:, 0].tolist(),
          "al_sum": float(np.abs(al).sum())}))
    ctx.close()
""") % ROOT


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_fitc_matches_reference_algorithm(world, tmp_path):
    """Data-sharded FITC (all-reduce of the M x M partial) against the CPU oracle on the full data set."""
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "wf.py"
    script.write_text(FITC_WORKER)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(29580 + world))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    res = {}
    for o in outs:
        for line in o[0].splitlines():
            if line.startswith("RESULT"):
                _, rank, payload = line.split(" ", 2)
                res[int(rank)] = json.loads(payload)
    from oracle import gp_oracle as go
    N, M, D = 3000, 150, 4
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, D)); y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    U = rng.standard_normal((M, D))
    spec = ("rbf", [np.log(2.0), 0.0])
    rpost, rnlZ, rdn = go.fitc_evaluate(("zero",), spec, U, np.log(0.1), X, y, 3)
    Xs = np.random.default_rng(1).standard_normal((40, D))
    rym, rys2, rfm, rfs2, _ = go.predict(("zero",), spec, np.log(0.1), X, rpost, Xs, xu=U)

    def rel(a, b):
        a = np.asarray(a, float).ravel(); b = np.asarray(b, float).ravel()
        return np.max(np.abs(a - b)) / np.max(np.abs(b))
    al_total = 0.0
    for rank in range(world):
        g = res[rank]
        assert abs(g["nlZ"] - rnlZ) < 1e-8 * abs(rnlZ), (world, rank, g["nlZ"], rnlZ)
        assert rel(g["alpha"], rpost["alpha"]) < 1e-5
        assert rel(g["dcov"], rdn["cov"]) < 1e-5 and abs(g["dlik"] - rdn["lik"][0]) < 1e-5 * abs(rdn["lik"][0])
        assert abs(g["Ltrace"] - np.trace(rpost["L"])) < 1e-4 * abs(np.trace(rpost["L"]))
        assert rel(g["ka"], rfm) < 1e-5 and rel(g["fs2"], rfs2) < 1e-5
        al_total += g["al_sum"]
    assert al_total > 0
