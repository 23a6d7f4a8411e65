"""world_size-2 gloo test (CPU) of the replica plumbing bench.py uses for --gpus N."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import sys, json
    sys.path.insert(0, %r)
    from pygps_b200._dist import DistCtx, replica_hyp, aggregate_rate
    ctx = DistCtx(backend="gloo")
    ctx.barrier()
    mine = 1.0 + ctx.rank            # pretend rank r took 1+r seconds
    tmax = ctx.max(mine)
    tsum = ctx.sum(mine)
    hyp = [replica_hyp(s, ctx.rank) for s in range(3)]
    ctx.barrier()
    if ctx.rank == 0:
        print(json.dumps({"max": tmax, "sum": tsum, "world": ctx.world,
                          "rate": aggregate_rate(10, ctx.world, tmax)}))
    print("HYP", ctx.rank, json.dumps(hyp))
    ctx.close()
""") % ROOT


def test_two_rank_gloo_barrier_and_max(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    line = [l for l in outs[0][0].splitlines() if l.startswith("{")][0]
    res = json.loads(line)
    assert res == {"max": 2.0, "sum": 3.0, "world": 2, "rate": 10.0}
    hyps = [l.split(" ", 2)[2] for o in outs for l in o[0].splitlines() if l.startswith("HYP")]
    assert len(hyps) == 2 and hyps[0] != hyps[1]            # replicas evaluate different hyper-parameters


def test_replica_hyp_all_distinct():
    from pygps_b200._dist import replica_hyp
    seen = {(tuple(h), s) for h, s in (replica_hyp(k, r) for k in range(50) for r in range(8))}
    assert len(seen) == 400
