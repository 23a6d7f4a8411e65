#!/usr/bin/env python
"""bench.py - nlZ evaluations/s of the exact-GP hot path at BASELINE.json's C2 configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full evaluation of what `GPR.getPosterior(der=False)` does
(K build, /sn2 + I, Cholesky, both triangular solves, log-det, nlZ back on the host;
/root/reference/pyGPs/Core/inf.py:353-370) with hyper-parameters that change every step, so
only X and y can stay cached.  Workload: N=16384, D=8, cov.RBF(log 2, 0), lik.Gauss(log 0.1),
mean.Zero, `default_rng(0)` synthetic data (reference nlZ 60824.3036486822).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  value     device-resident rate: X resident in HBM, per step only (y-m) goes up and alpha/nlZ come back
  e2e       the same metric through the plugin API `model.getPosterior(x, y, der=False)` with HOST
            arrays: X and y are uploaded and alpha/nlZ downloaded inside the timed region, every step
  roofline  trailing SYRK update (oz_syrk_kernel, int8 tensor cores): algorithmic fp64 flops kw*n_t^2 per step
            x 28 int8 products, over CUDA-event time on the launching stream, against the int8 tensor peak; the
            fp64-equivalent rate is given beside the fp64 pipe peak measured
            on this box by the library's own DMMA micro-benchmark (MEASURED_PEAKS.json has no fp64 figure)
  cpu_baseline  the UNMODIFIED reference (pyGPs.GPR().getPosterior from baseline/_ref, kind "reference"; the pinned
            oracle port when that install is absent, kind "port") on the host cores, on a bounded sample
  sharded   (N > 1 only) ONE evaluation of the C3 family sharded over the N GPUs (gpk_exact_eval_dist: block-cyclic
            columns, NCCL panel broadcasts) and one sharded FITC evaluation (C4 family, all-reduce), with parity
            against the single-GPU path - the replicas above have no collective, this block is what exercises NCCL
  other_configs  (N = 1) C1 latency, C3 on one GPU, C4, C5 and the der=True rate, timed in the same driver run
`--impl reference` runs the unmodified reference ONCE (twice if it is fast) at the FULL configuration - a real
N=16384 evaluation, 1-2 minutes of host time - and reports the measured time; `steps`/`warmup` in its line are
the ones actually executed.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = None   # the original stdout (see main)
C3_N, C3_D = 65536, 32
C4_N, C4_M, C4_D = 262144, 4096, 8
C5_N, C5_D = 8192, 16
METRIC = "nlZ evals/sec at N=16384 D=8 RBF"
UNIT = "evals/s"
N_FULL, D_FULL = 16384, 8
NB = 128


def synth(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    return X, y


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi sampling DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Index of the next sample: call at the start and at the end of the timed region."""
        return len(self.lines)

    def stop(self, lo=0, hi=None):
        """Summary of the samples [lo, hi) (the timed region).  The sampler is started before the warm-up (nvidia-smi
        needs a few hundred ms to deliver its first line, the timed region is ~0.25 s); if the region is too short to
        hold a sample, the closest samples - taken under the same workload during the warm-up - are used and flagged."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        hi = len(self.lines) if hi is None else hi
        window, note = self.lines[lo:hi], None
        if not window:
            window = self.lines[max(0, lo - 5):hi + 1]
            note = "timed region shorter than the sampling interval: samples from the warm-up of the same workload"
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self._note = note
        for l in window:
            f = [s.strip() for s in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
               "samples": len(sm), "reasons": sorted(reasons)}
        if self._note:
            out["note"] = self._note
        return out


# ----------------------------------------------------------------------------- shared by both arms
def make_config(N, D, n_gpus):
    """The `config` object - identical for the GPU arm and the reference arm (the driver compares them)."""
    return {"workload": "C2: GPR Exact, cov.RBF, N=%d D=%d fp64 (K build + Cholesky + solves + nlZ)" % (N, D),
            "parallelism": "replicas x%d (independent hyper-parameter vectors per GPU, no collective)" % n_gpus,
            "l2": "working set (%.1f GiB factor) exceeds the 126 MB L2; no flush needed" % (8.0 * N * N / 2 ** 30),
            "hyp": "changed every step"}


# ----------------------------------------------------------------------------- CPU legs
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline must get every host core it can use."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def _blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def reference_evaluator():
    """(kind, fn) with fn(X, y, hyp, log_sn) -> nlZ running the reference's CPU path for one evaluation.

    kind "reference": the UNMODIFIED reference package installed into baseline/_ref (git-ignored, travels to the GPU
    box), imported through the three stub modules of oracle/shim (past.utils, past.builtins, matplotlib.pyplot) and
    driven through its own public API, pyGPs.GPR().getPosterior(x, y, der=False).
    kind "port": the oracle's pinned restatement of the same algorithm (only when baseline/_ref is absent)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "pyGPs")):
        try:
            for pth in (os.path.join(ROOT, "oracle", "shim"), ref_dir):
                if pth not in sys.path:
                    sys.path.insert(0, pth)
            import logging
            import warnings
            warnings.filterwarnings("ignore")
            import pyGPs
            logging.disable(logging.WARNING)

            def run(X, y, hyp, log_sn):
                m = pyGPs.GPR()
                m.setPrior(kernel=pyGPs.cov.RBF(hyp[0], hyp[1]))
                m.setNoise(log_sn)
                nlZ, post = m.getPosterior(X, y, der=False)
                return float(nlZ)
            return "reference", run
        except Exception as e:          # pragma: no cover
            sys.stderr.write("baseline/_ref present but not importable (%s); using the oracle port\n" % e)
    from oracle import gp_oracle as go

    def run_port(X, y, hyp, log_sn):
        return float(go.exact_evaluate(("zero",), ("rbf", list(hyp)), log_sn, X, y, 2)[1])
    return "port", run_port


def cpu_sample(n_full, d, n_sample, full_budget_s=90.0):
    """cpu_baseline of the GPU arm.  The reference's time does NOT scale smoothly with N (measured here and on the GPU
    box: N=8192 takes 13-23 s, N=16384 only 25-60 s - its two general `solve` calls hit different BLAS regimes), so a
    scaled sample misjudges it by up to 3.4x.  Hence: a 2-5 s probe at N=4096, and if the full size is predicted to fit
    `full_budget_s` (it takes ~25 s on the GPU box's 16 cores) ONE FULL-SIZE evaluation is measured; only otherwise
    the N_s sample scaled with the exponent measured between N_s/2 and N_s."""
    from pygps_b200._dist import replica_hyp
    _use_all_host_threads()
    kind, fn = reference_evaluator()
    what = "unmodified reference pyGPs.GPR().getPosterior(der=False)" if kind == "reference" else "oracle port"
    h, sn = replica_hyp(0, 0)

    def run(n):
        X, y = synth(n, d)
        t0 = time.perf_counter()
        fn(X, y, h, sn)
        return time.perf_counter() - t0
    t = {}
    probe = min(4096, n_full)
    t[probe] = run(probe)
    if n_full == probe or t[probe] * (n_full / float(probe)) ** 2 <= full_budget_s:
        if n_full != probe:
            t[n_full] = run(n_full)
        sample = "1 FULL-SIZE evaluation of the %s at N=%d D=%d: %.2f s (after a %.2f s probe at N=%d)" % (
            what, n_full, d, t[n_full], t[probe], probe)
        return {"value": 1.0 / t[n_full], "unit": UNIT, "cores": _blas_threads(), "kind": kind, "sample": sample}
    for n in (n_sample // 2, n_sample):
        if n not in t:
            t[n] = run(n)
    p = min(3.0, max(2.0, math.log(t[n_sample] / t[n_sample // 2], 2.0)))
    per_full = t[n_sample] * (n_full / float(n_sample)) ** p
    sample = ("1 evaluation of the %s at N=%d D=%d: %.2f s; scaled to N=%d by (N/N_s)^p, p=%.2f measured here between "
              "N=%d and N=%d - an estimate only: the full-size measurement is `bench.py --impl reference`"
              % (what, n_sample, d, t[n_sample], n_full, p, n_sample // 2, n_sample))
    return {"value": 1.0 / per_full, "unit": UNIT, "cores": _blas_threads(), "kind": kind, "sample": sample}


def run_reference(args):
    """--impl reference: the unmodified reference on this box's host cores, at the FULL configuration, measured."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from pygps_b200._dist import replica_hyp
    _use_all_host_threads()
    kind, fn = reference_evaluator()
    N, D = args.n, args.d
    # warm the BLAS threads / page the libraries in on a small problem (not a timed step, not counted)
    Xw, yw = synth(min(N, 2048), D)
    fn(Xw, yw, [math.log(2.0), 0.0], math.log(0.1))
    X, y = synth(N, D)
    times, nlz = [], None
    budget_s = float(os.environ.get("GPK_BENCH_REF_BUDGET_S", "240"))
    max_steps = max(1, min(args.steps, 2))
    for k in range(max_steps):
        h, sn = ([math.log(2.0), 0.0], math.log(0.1)) if k == 0 else replica_hyp(k, 0)
        t0 = time.perf_counter()
        v = fn(X, y, h, sn)
        times.append(time.perf_counter() - t0)
        if k == 0:
            nlz = v
        if sum(times) + times[-1] > budget_s:
            break
    per = sum(times) / len(times)
    ref_nlz = 60824.3036486822 if (N, D) == (16384, 8) else None
    what = "unmodified reference (baseline/_ref) pyGPs.GPR().getPosterior(der=False)" if kind == "reference" \
        else "oracle port of the reference's numpy/scipy path"
    line = {"impl": "reference", "metric": METRIC, "value": 1.0 / per, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": 0, "ms_per_step": 1e3 * per, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(N, D, args.gpus),
            "engine": what + ", host CPU, %d BLAS threads" % _blas_threads(),
            "requested": {"steps": args.steps, "warmup": args.warmup,
                          "note": "a full-size evaluation takes 1-2 min of host time: 1-2 measured steps are run, "
                                  "never an extrapolation; one untimed N=2048 call warms the BLAS threads"},
            "cpu_baseline": {"value": 1.0 / per, "unit": UNIT, "cores": _blas_threads(), "kind": kind,
                             "sample": "%d full evaluation(s) at N=%d D=%d, %s s each" % (
                                 len(times), N, D, ", ".join("%.1f" % t for t in times))},
            "e2e": {"value": 1.0 / per, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "parity": {"nlZ": nlz, "reference_nlZ": ref_nlz,
                       "rel_err": None if ref_nlz is None else abs(nlz - ref_nlz) / abs(ref_nlz)},
            "gpu_launches": 0}
    _emit(line)
    return 0


def _emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _residual(X, y, alpha, ell, sn2, rows=256):
    """|(K + sn2 I) alpha - y| / |y| on sampled rows (host numpy; self-consistency where no oracle can run)."""
    N = X.shape[0]
    idx = np.random.default_rng(1).choice(N, size=min(rows, N), replace=False)
    Xs = X / ell
    d2 = np.sum(Xs[idx] ** 2, 1)[:, None] + np.sum(Xs ** 2, 1)[None, :] - 2 * Xs[idx] @ Xs.T
    res = np.exp(-0.5 * np.maximum(d2, 0)) @ alpha + sn2 * alpha[idx] - y[idx]
    return float(np.linalg.norm(res) / np.linalg.norm(y[idx]))


# ----------------------------------------------------------------------------- sharded evaluations (N > 1)
def run_sharded(ctx, args, eng_single):
    """ONE evaluation sharded over all ranks: C3 family through gpk_exact_eval_dist (NCCL panel broadcasts) and the C4
    family through the data-sharded gpk_fitc_eval (NCCL all-reduce), each checked against the single-GPU path."""
    from pygps_b200 import _lib
    out = {"n_gpus": ctx.world}
    eng = _lib.Engine(ctx.local_rank)
    ctx.shard_engine(eng)
    # ---- C3: GPR Exact, cov.RBFard, N=65536, D=32 ------------------------------------------------------------
    N, D = args.sharded_n, C3_D
    X, y = synth(N, D)
    yv = y.reshape(-1)
    hyp = [math.log(3.0)] * D + [0.0]
    eng.set_data(X)
    ts, st = [], None
    for i in range(3):
        ctx.barrier()
        nlZ, alpha = eng.exact_eval_dist(_lib.COV_RBFARD, 3, hyp, math.log(0.1), yv)
        st = eng.stats()
        ts.append(ctx.max(st["total_ms"]))
    potrf_ms = ctx.max(st["potrf_ms"])
    best = min(ts[1:])
    c3 = {"workload": "C3 family: GPR Exact, cov.RBFard, N=%d D=%d fp64, block-cyclic columns over %d GPUs, NCCL panel "
                      "broadcast per step" % (N, D, ctx.world),
          "ms_per_eval": best, "evals_per_s": 1e3 / best, "cholesky_tflops": N ** 3 / 3.0 / (potrf_ms * 1e-3) / 1e12,
          "nlZ": float(nlZ), "int8_trailing_update": os.environ.get("GPK_DIST_OZAKI", "1") != "0"}
    if ctx.rank == 0:
        c3["residual"] = _residual(X, y, alpha, 3.0, 0.01)
        eng_single.set_data(X)
        nl1, a1, _, _ = eng_single.exact_eval(_lib.COV_RBFARD, 3, hyp, math.log(0.1), yv, False)
        nl1, a1, _, _ = eng_single.exact_eval(_lib.COV_RBFARD, 3, hyp, math.log(0.1), yv, False)
        c3["single_gpu_ms"] = eng_single.stats()["total_ms"]
        c3["rel_err_vs_single_gpu"] = {"nlZ": abs(float(nlZ) - float(nl1)) / abs(float(nl1)),
                                       "alpha": float(np.max(np.abs(alpha - a1)) / np.max(np.abs(a1)))}
        if not (c3["rel_err_vs_single_gpu"]["nlZ"] < 1e-10 and c3["residual"] < 1e-8):
            raise SystemExit("parity gate failed for the sharded evaluation: %r" % c3)
    out["c3"] = c3
    ctx.barrier()
    del X
    # ---- C2 data, sharded WITH derivatives and predictions: factor gather + row-slab inverse ----
    N2, D2 = N_FULL, D_FULL
    X2, y2 = synth(N2, D2)
    h2, sn2l = [math.log(2.0), 0.0], math.log(0.1)
    Xs2 = np.random.default_rng(1).standard_normal((4096, D2))
    lo, hi = (4096 * ctx.rank) // ctx.world, (4096 * (ctx.rank + 1)) // ctx.world
    eng.set_data(X2)
    eng.dist_reserve(2)
    ts = []
    for i in range(2):
        ctx.barrier()
        t0 = time.perf_counter()
        nl, al, dc, dl = eng.exact_eval_dist_der(_lib.COV_RBF, 3, h2, sn2l, y2.reshape(-1))
        ts.append(ctx.max(1e3 * (time.perf_counter() - t0)))
    ctx.barrier()
    t0 = time.perf_counter()
    eng.dist_gather_factor()
    ka, fs2 = eng.predict(Xs2[lo:hi]) if hi > lo else (np.empty((0, 1)), np.empty((0, 1)))
    t_pred = ctx.max(1e3 * (time.perf_counter() - t0))
    e4 = {"workload": "C2 data (N=%d D=%d) sharded over %d GPUs: evaluation WITH derivatives (factor gather, row-slab "
                      "inverse, one all-reduce) and 4096 predictions split over the ranks" % (N2, D2, ctx.world),
          "der_eval_ms": min(ts), "predict_ms": t_pred, "nlZ": float(nl)}
    if ctx.rank == 0:
        eng_single.set_data(X2)
        r1 = eng_single.exact_eval(_lib.COV_RBF, 3, h2, sn2l, y2.reshape(-1), True)
        ka1, fs21 = eng_single.predict(Xs2[lo:hi])
        rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))
        e4["rel_err_vs_single_gpu"] = {"nlZ": abs(float(nl) - float(r1[0])) / abs(float(r1[0])), "dcov": rel(dc, r1[2]),
                                       "dlik": rel(dl, r1[3]), "predict_mean": rel(ka, ka1), "predict_var": rel(fs2, fs21)}
        if not max(e4["rel_err_vs_single_gpu"].values()) < 1e-8:
            raise SystemExit("parity gate failed for the sharded derivatives / predictions: %r" % e4)
    out["sharded_der_predict"] = e4
    ctx.barrier()
    del X2
    # ---- C4: GPR_FITC, cov.RBF, N=262144, M=4096 inducing points, data sharded ----------------------------------
    N4, M4 = args.fitc_n, args.fitc_m
    rng = np.random.default_rng(0)
    X4 = rng.standard_normal((N4, C4_D))
    y4 = np.sin(X4.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N4, 1))
    U4 = rng.standard_normal((M4, C4_D))
    lo, hi = (N4 * ctx.rank) // ctx.world, (N4 * (ctx.rank + 1)) // ctx.world
    eng.set_data(X4[lo:hi])
    ts = []
    for i in range(3):
        ctx.barrier()
        f = eng.fitc_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), U4, y4[lo:hi].reshape(-1), False)
        ts.append(ctx.max(eng.stats()["total_ms"]))
    c4 = {"workload": "C4 family: GPR_FITC, cov.RBF, N=%d M=%d D=%d fp64, data points sharded over %d GPUs, NCCL "
                      "all-reduce of the M x M partial" % (N4, M4, C4_D, ctx.world),
          "ms_per_eval": min(ts[1:]), "nlZ": float(f[0])}
    if ctx.rank == 0:
        eng_single.set_data(X4)
        g = eng_single.fitc_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), U4, y4.reshape(-1), False)
        g = eng_single.fitc_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), U4, y4.reshape(-1), False)
        c4["single_gpu_ms"] = eng_single.stats()["total_ms"]
        c4["rel_err_vs_single_gpu"] = {"nlZ": abs(float(f[0]) - float(g[0])) / abs(float(g[0])),
                                       "alpha": float(np.max(np.abs(f[1] - g[1])) / np.max(np.abs(g[1])))}
        if not c4["rel_err_vs_single_gpu"]["nlZ"] < 1e-9:
            raise SystemExit("parity gate failed for the sharded FITC evaluation: %r" % c4)
    out["c4"] = c4
    ctx.barrier()
    return out


def run_concurrent(device, X, ymm, streams, evals):
    """Throughput of `streams` concurrent evaluation streams on ONE GPU (one libgpk handle + one thread each, independent
    hyper-parameter vectors): the chain-bound tail of one factorisation overlaps the bulk of another.  Reported beside
    the headline, which stays the single-stream rate."""
    import threading
    from pygps_b200 import _lib
    from pygps_b200._dist import replica_hyp
    engs = [_lib.Engine(device) for _ in range(streams)]
    for e in engs:
        e.set_data(X)
        for k in range(2):
            e.exact_eval(_lib.COV_RBF, 3, *replica_hyp(k, 7), ymm, False)
    per = max(1, evals // streams)

    def work(i):
        for k in range(per):
            h, sn = replica_hyp(50 + k * streams + i, 7)
            engs[i].exact_eval(_lib.COV_RBF, 3, h, sn, ymm, False)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(streams)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return {"streams": streams, "evals": per * streams, "value": per * streams / dt, "unit": UNIT,
            "note": "NOT the headline: `value` above is one evaluation at a time"}


# ----------------------------------------------------------------------------- other configs on one GPU (N = 1)
def run_other_configs(eng, args):
    from pygps_b200 import _lib
    out = {}
    # C1: N=512, D=2 - per-evaluation latency floor (the reference's own CPU-runnable case)
    rng = np.random.default_rng(0)
    X1 = rng.standard_normal((512, 2)); y1 = np.sin(X1.sum(1)) + 0.1 * rng.standard_normal(512)
    eng.set_data(X1)
    ts = []
    for k in range(8):
        t0 = time.perf_counter()
        eng.exact_eval(_lib.COV_RBF, 3, [0.01 * k, 0.0], math.log(0.1), y1, False)
        ts.append(1e3 * (time.perf_counter() - t0))
    out["c1_n512_d2"] = {"latency_ms_wall": min(ts[2:]), "device_ms": eng.stats()["total_ms"]}
    # C3 on ONE GPU (int8 trailing update)
    X, y = synth(C3_N, C3_D)
    hyp = [math.log(3.0)] * C3_D + [0.0]
    eng.set_data(X)
    best = None
    for i in range(2):
        nlZ, alpha, _, _ = eng.exact_eval(_lib.COV_RBFARD, 3, hyp, math.log(0.1), y.reshape(-1), False)
        st = eng.stats()
        if best is None or st["total_ms"] < best["total_ms"]:
            best = st
    out["c3_single_gpu"] = {"workload": "GPR Exact, cov.RBFard, N=%d D=%d fp64, 1 GPU" % (C3_N, C3_D),
                            "ms_per_eval": best["total_ms"],
                            "cholesky_tflops_fp64_equivalent": C3_N ** 3 / 3.0 / (best["potrf_ms"] * 1e-3) / 1e12,
                            "nlZ": float(nlZ), "residual": _residual(X, y, alpha, 3.0, 0.01)}
    del X
    # C4 at full size on one GPU
    rng = np.random.default_rng(0)
    X4 = rng.standard_normal((C4_N, C4_D))
    y4 = np.sin(X4.sum(1)) + 0.1 * rng.standard_normal(C4_N)
    U4 = rng.standard_normal((C4_M, C4_D))
    eng.set_data(X4)
    ts = []
    for i in range(3):
        f = eng.fitc_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), U4, y4, False)
        ts.append(eng.stats()["total_ms"])
    out["c4_fitc"] = {"workload": "GPR_FITC, cov.RBF, N=%d M=%d fp64, 1 GPU" % (C4_N, C4_M),
                      "ms_per_eval": min(ts[1:]), "nlZ": float(f[0])}
    del X4
    # C5 at full size (GPC / EP)
    rng = np.random.default_rng(0)
    X5 = rng.standard_normal((C5_N, C5_D))
    lab = np.sign(X5[:, 0] + 0.5 * X5[:, 1] + 0.3 * rng.standard_normal(C5_N))
    lab[lab == 0] = 1
    eng.set_data(X5)
    ts = []
    for i in range(2):
        t0 = time.perf_counter()
        r = eng.ep_eval(_lib.COV_RBF, 3, [math.log(4.0), 0.0], np.zeros(C5_N), lab, None, None, False, True)
        ts.append(1e3 * (time.perf_counter() - t0))
    out["c5_ep"] = {"workload": "GPC EP, cov.RBF, N=%d D=%d fp64, 1 GPU, incl. derivatives" % (C5_N, C5_D),
                    "ms_per_eval": min(ts), "sweeps": int(r[7]), "nlZ": float(r[0])}
    return out


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    from pygps_b200 import _lib, build
    from pygps_b200._dist import DistCtx, replica_hyp, aggregate_rate
    import pygps_b200 as pg

    ctx = DistCtx()
    build.build()
    if ctx.world != args.gpus and ctx.rank == 0 and ctx.world > 1:
        sys.stderr.write("note: WORLD_SIZE=%d differs from --gpus %d; using WORLD_SIZE\n" % (ctx.world, args.gpus))
    n_gpus = ctx.world
    eng = _lib.Engine(ctx.local_rank)
    N, D = args.n, args.d
    X, y = synth(N, D)
    pin = None
    try:
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(ctx.local_rank)
            pin = (torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory())
            X, y = pin[0].numpy(), pin[1].numpy()
    except Exception:
        pin = None
    ymm = y.reshape(-1)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    clocks = ClockSampler(ctx.local_rank)

    # ---- device-resident arm ------------------------------------------------------------
    eng.set_data(X)
    # two timing events around each level-1 trailing update on its own stream: no synchronisation, no added
    # dependency - the schedule is the one an unprofiled evaluation runs (include/gpk.h: gpk_set_profile)
    eng.set_profile(True)
    nlz_first = None
    clocks.start()             # before the warm-up: nvidia-smi's first line takes a few hundred ms
    for k in range(warmup):
        h, sn = replica_hyp(k, ctx.rank)
        out = eng.exact_eval(_lib.COV_RBF, 3, h, sn, ymm, False)
    # parity gate before any number is reported: the reference's hyper-parameters, the reference's nlZ
    nlz_first = eng.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), ymm, False)[0]
    ref_nlz = 60824.3036486822 if (N, D) == (16384, 8) else None
    parity = None if ref_nlz is None else abs(nlz_first - ref_nlz) / abs(ref_nlz)
    if parity is not None and not parity < 1e-6:
        raise SystemExit("parity gate failed: nlZ %r vs reference %r" % (nlz_first, ref_nlz))
    ctx.barrier()
    clk_lo = clocks.mark()
    dev_ms = 0.0
    launches = 0
    stage = {"kbuild_ms": 0.0, "potrf_ms": 0.0, "solve_ms": 0.0}
    syrk_ms = syrk_fl = 0.0
    t0 = time.perf_counter()
    for k in range(steps):
        h, sn = replica_hyp(warmup + k, ctx.rank)
        eng.exact_eval(_lib.COV_RBF, 3, h, sn, ymm, False)
        st = eng.stats()
        dev_ms += st["total_ms"]
        launches += st["launches"]
        for key in stage:
            stage[key] += st[key]
        syrk_ms += st["syrk_ms"]
        syrk_fl += st["syrk_flops"]
    wall = time.perf_counter() - t0            # every call ends with a stream synchronize
    clk_hi = clocks.mark()
    ctx.barrier()
    t_rank = max(wall, dev_ms * 1e-3)
    t_max = ctx.max(t_rank)
    value = aggregate_rate(steps, n_gpus, t_max)
    clk = clocks.stop(clk_lo, clk_hi)
    launches_total = int(ctx.sum(launches))

    # ---- end-to-end arm through the plugin API (host arrays in, host results out) ----------
    model = pg.GPR()
    model.inffunc._engine = eng
    e2e_h2d = X.nbytes + y.nbytes + 3 * 8
    e2e_d2h = N * 8 + 8 * 8 + 4
    for k in range(2):
        h, sn = replica_hyp(1000 + k, ctx.rank)
        model.covfunc.hyp = h; model.likfunc.hyp = [sn]
        model.getPosterior(X, y, der=False)
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        h, sn = replica_hyp(2000 + k, ctx.rank)
        model.covfunc.hyp = h; model.likfunc.hyp = [sn]
        nlz_e2e, post = model.getPosterior(X, y, der=False)
    e2e_wall = time.perf_counter() - t0
    ctx.barrier()
    e2e_value = aggregate_rate(steps, n_gpus, ctx.max(e2e_wall))

    # ---- derivative rate (what optimize() sees), reported beside the headline ---------------
    der_ms = None
    if ctx.rank == 0 and not args.no_der:
        eng.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), ymm, True)
        t0 = time.perf_counter()
        eng.exact_eval(_lib.COV_RBF, 3, [math.log(2.05), 0.01], math.log(0.1), ymm, True)
        der_ms = 1e3 * (time.perf_counter() - t0)
    ctx.barrier()

    # ---- roofline of the dominant kernel (rank 0; measured live inside the timed region above) ----
    roof = None
    extra = {}
    if ctx.rank == 0:
        eng.set_profile(False)
        reps = steps
        peaks = {}
        for shape, name in ((0, "m8n8k4"), (2, "m16n8k8"), (3, "m16n8k16"), (4, "dfma")):
            try:
                peaks[name] = max(eng.bench_dmma(shape, w, 4000)[0] for w in (4, 8, 16))
            except Exception as e:           # pragma: no cover
                peaks[name] = None
        fp64_peak = max(v for k, v in peaks.items() if v and k != "dfma")
        fp64_equiv = syrk_fl / (syrk_ms * 1e-3) / 1e12          # algorithmic fp64 flops of the level-1 updates / their time
        oz_on = os.environ.get("GPK_OZAKI", "1") != "0"
        iso_ms, iso_tf = eng.bench_syrk(N - NB, NB, 5)
        measured = {}
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
                measured = json.load(f)
        except Exception:
            measured = {}
        if oz_on:
            # The trailing update runs on the int8 tensor pipe: every fp64 multiply-add of the update is S(S+1)/2 = 28
            # exact int8 multiply-adds (7 radix-256 slices per operand, products with t+u <= 8).  Roofline = executed
            # int8 tensor ops against TWO denominators: (a) 2x the measured bf16 dense figure of MEASURED_PEAKS.json
            # (same pipe, 32 instead of 16 K-elements per instruction) - the contract's `peak`; (b) the int8 ISSUE-RATE
            # peak at the SM clock sampled during the timed region (16384 ops/clk/SM: one 128x256x32 MMA per 128 clk,
            # measured by gpk_bench_i8_rate) - the stricter one, since this kernel is not power-bound like a bf16 GEMM.
            PRODUCTS = 28
            achieved = PRODUCTS * fp64_equiv
            try:
                probe_clk, probe_tops = eng.bench_i8_rate(256, 40000, 4096, 128, 0, 0, 148)
            except Exception:            # pragma: no cover
                probe_clk, probe_tops = None, None
            if measured.get("bf16_tflops_sustained"):
                peak = 2.0 * measured["bf16_tflops_sustained"]
                src = "2 x MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
            else:
                peak = 2.0 * 1590.0
                src = "2 x the profiling guide's fallback bf16 figure (1.59 PFLOP/s); MEASURED_PEAKS.json absent: of fallback"
            issue_peak = None
            if probe_clk and clk.get("sm_mhz"):
                issue_peak = 148 * (2.0 * 128 * 256 * 32 / probe_clk) * clk["sm_mhz"] * 1e6 / 1e12
            traffic, traffic_note = None, None
            for name in ("r2_oz_syrk_traffic.json", "r1_oz_syrk_traffic.json"):
                try:
                    with open(os.path.join(ROOT, "profiles", name)) as f:
                        tj = json.load(f)
                    traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                    traffic_note = {"launch": tj["launch"], "algorithmic_bytes": tj["algorithmic_bytes"]["total"],
                                    "ratio": traffic / float(tj["algorithmic_bytes"]["total"]), "source": tj["source"]}
                    break
                except Exception:
                    pass
            roof = {"bound": "tensor", "kernel": "oz_syrk_kernel<7,8,1,0> (trailing SYRK update: tcgen05.mma.kind::i8, TMEM accumulators)",
                    "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_note": traffic_note,
                    "peak_source": src,
                    "int8_issue_peak": {"tops_at_sampled_clock": issue_peak, "sm_mhz": clk.get("sm_mhz"),
                                        "frac": None if not issue_peak else achieved / issue_peak,
                                        "clk_per_mma_128x256x32": probe_clk, "probe_tops": probe_tops,
                                        "note": "in-situ average over every level-1 update of the timed steps, SMs "
                                                "shared with the panel streams; profiles/ holds the ncu figure of the "
                                                "largest launch in isolation"},
                    "fp64_equivalent": {"achieved_tflops": fp64_equiv, "fp64_pipe_peak_tflops": fp64_peak,
                                        "ratio_to_fp64_pipe_peak": fp64_equiv / fp64_peak,
                                        "int8_products_per_fp64_mac": PRODUCTS},
                    "flops_per_eval": syrk_fl / reps}
        else:
            roof = {"bound": "tensor", "kernel": "dgemm_nt_kernel<1> (trailing SYRK/GEMM update, fp64 DMMA)",
                    "achieved": fp64_equiv, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_equiv / fp64_peak,
                    "traffic": None,
                    "peak_source": "fp64 DMMA micro-benchmark gpk_bench_dmma on this box (MEASURED_PEAKS.json has no fp64 figure)",
                    "flops_per_eval": syrk_fl / reps}
        extra = {"fp64_peaks_tflops": peaks, "syrk_isolated_dmma": {"n": N - NB, "k": NB, "ms": iso_ms, "tflops": iso_tf},
                 "hbm_copy_gbs": eng.bench_copy(1 << 30, 5),
                 "cholesky_tflops_in_eval": (N ** 3 / 3.0) / (stage["potrf_ms"] / steps * 1e-3) / 1e12,
                 "stage_ms_per_eval": {k: v / steps for k, v in stage.items()},
                 "der_eval_ms": der_ms}

    # ---- several evaluation streams on the one GPU (independent evaluations, as random restarts are) ----------
    if ctx.rank == 0 and n_gpus == 1 and not args.no_extra:
        try:
            extra["concurrent_streams"] = run_concurrent(ctx.local_rank, X, ymm, 2, max(steps, 10))
        except Exception as e:           # pragma: no cover
            extra["concurrent_streams"] = {"error": repr(e)}

    # ---- one evaluation SHARDED over all GPUs (N > 1): the NCCL paths, with parity ---------------
    sharded = None
    if n_gpus > 1 and not args.no_sharded:
        sharded = run_sharded(ctx, args, eng)

    # ---- the other BASELINE configs on one GPU, in the same driver run (N = 1) -------------------
    if ctx.rank == 0 and n_gpus == 1 and not args.no_extra:
        try:
            extra["other_configs"] = run_other_configs(eng, args)
        except Exception as e:           # pragma: no cover
            extra["other_configs"] = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N=1 only) ------------------------------------------
    cpu = None
    if ctx.rank == 0 and n_gpus == 1 and not args.no_cpu:
        cpu = cpu_sample(N, D, 8192 if N >= 8192 else N)
        # also a "fair CPU" line (NOT the reference: in-place cdist+exp, cho_factor, cho_solve), so the
        # speed-up is not credited for the reference's LU-on-a-triangular-matrix waste.
        try:
            from oracle import gp_oracle as go
            nf = 8192 if N >= 8192 else N
            Xf, yf = synth(nf, D)
            t0f = time.perf_counter()
            go.exact_evaluate_fair(("rbf", [math.log(2.0), 0.0]), math.log(0.1), Xf, yf)
            tf = time.perf_counter() - t0f
            cpu["fair_cpu"] = {"value": 1.0 / (tf * (N / float(nf)) ** 3), "unit": UNIT,
                               "sample": "scipy cho_factor/cho_solve path at N=%d in %.2f s, scaled by (N/N_s)^3" % (nf, tf)}
        except Exception as e:          # pragma: no cover
            cpu["fair_cpu"] = {"error": str(e)}

    if ctx.rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": 1e3 * t_max / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": make_config(N, D, n_gpus),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                        "api": "pygps_b200.GPR().getPosterior(x, y, der=False) with pinned host arrays"},
                "gpu_launches": launches_total, "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
                "parity": {"nlZ": nlz_first, "reference_nlZ": ref_nlz, "rel_err": parity}}
        if sharded is not None:
            line["sharded"] = sharded
        line.update(extra)
        _emit(line)
    ctx.close()
    return 0


def main():
    # rank 0 prints ONE JSON line on stdout.  NCCL prints its version / debug lines on the process's stdout (fd 1) when
    # the box exports NCCL_DEBUG, so fd 1 is pointed at stderr for the whole run and the JSON line goes to a private
    # duplicate of the original stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problem-n", dest="n", type=int, default=N_FULL)
    ap.add_argument("--problem-d", dest="d", type=int, default=D_FULL)
    ap.add_argument("--sharded-n", dest="sharded_n", type=int, default=C3_N, help="N of the sharded C3-family evaluation")
    ap.add_argument("--fitc-n", dest="fitc_n", type=int, default=C4_N)
    ap.add_argument("--fitc-m", dest="fitc_m", type=int, default=C4_M)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-der", action="store_true", help="skip the derivative-rate side measurement")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sharded C3/C4 evaluations")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the other BASELINE configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
