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
    for N, D in ((300, 3), (2048, 8), (4096, 8)):
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


def _run(world, tmp_path, port):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
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


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_eval_matches_single_gpu(world, tmp_path, golden):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    res = _run(world, tmp_path, 29560 + world)
    assert len(res) == world
    g = golden("synthetic")
    for N, D in ((300, 3), (2048, 8), (4096, 8)):
        ref = _reference(N, D)
        for rank in range(world):
            got = res[rank][str(N)]
            assert abs(got[0] - ref[0]) < 1e-10 * abs(ref[0]), (world, rank, N, got, ref)
            assert abs(got[1] - ref[1]) < 1e-8 * abs(ref[1]) and abs(got[2] - ref[2]) < 1e-7 * max(1, abs(ref[2]))
    assert abs(res[0]["2048"][0] - float(g["c2_2048_nlZ"])) < 1e-9 * abs(float(g["c2_2048_nlZ"]))
