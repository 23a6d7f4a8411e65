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
  cpu_baseline  the reference's algorithm (oracle port: cdist+exp, dpotrf, 2x dgesv) on the host cores
`--impl reference` times that same CPU port on this configuration (the reference is pure Python and
cannot travel to the GPU box; `oracle/` is its pinned restatement - see DESIGN.md).
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


def cpu_port_measure(n_full, d, per_step_budget_s, steps=1, warmup=0):
    """Time the oracle port (the reference's algorithm: cdist+exp, dpotrf on a Fortran copy, two dgesv on
    the triangular factor) on the host cores.

    The full size is run whenever it fits the per-step budget.  Otherwise the largest power-of-two
    sample N_s that fits is run and the result is scaled to N with the exponent p MEASURED on this box
    between N_s/2 and N_s (clipped to [2,3]) - the reference's time is sub-cubic in this range because
    BLAS efficiency grows with size, so a plain cubic scaling would flatter the GPU.
    Returns (evals/s at n_full, seconds per step at the sample size, sample description, cores)."""
    from oracle import gp_oracle as go
    from pygps_b200._dist import replica_hyp
    _use_all_host_threads()

    def one(X, y, k):
        h, sn = replica_hyp(k, 0)
        t = time.perf_counter()
        go.exact_evaluate(("zero",), ("rbf", h), sn, X, y, 2)
        return time.perf_counter() - t

    sizes = [n_full]
    while sizes[-1] > 1024:
        sizes.append(sizes[-1] // 2)
    sizes = sizes[::-1]                       # ascending: ..., n/4, n/2, n
    timings = {}
    n_s = sizes[0]
    for i, n in enumerate(sizes):
        if i >= 2:
            p_est = math.log(timings[sizes[i - 1]] / timings[sizes[i - 2]], 2.0)
            predicted = timings[sizes[i - 1]] * 2.0 ** min(3.0, max(2.0, p_est))
        elif i == 1:
            predicted = timings[sizes[0]] * 8.0
        else:
            predicted = 0.0
        if predicted > per_step_budget_s:
            break
        X, y = synth(n, d)
        timings[n] = one(X, y, 0)
        n_s = n
    X, y = synth(n_s, d)
    for k in range(warmup):
        one(X, y, 100 + k)
    t0 = time.perf_counter()
    for k in range(steps):
        one(X, y, 200 + k)
    per = (time.perf_counter() - t0) / steps
    sample = "%d eval(s) of the reference algorithm (oracle port) at N=%d D=%d, %.2f s each" % (steps, n_s, d, per)
    factor = 1.0
    if n_s != n_full:
        half = n_s // 2
        p = 3.0
        if half in timings:
            p = min(3.0, max(2.0, math.log(timings[n_s] / timings[half], 2.0)))
        factor = (n_full / float(n_s)) ** p
        sample += "; scaled to N=%d by (N/N_s)^p with p=%.2f measured here between N=%d and N=%d" % (
            n_full, p, half, n_s)
    return 1.0 / (per * factor), per, sample, _blas_threads()


def run_reference(args):
    """--impl reference: the reference's CPU path (pinned oracle port) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget = 170.0 / (steps + warmup + 1)      # whole run bounded to a few minutes
    value, per, sample, cores = cpu_port_measure(args.n, args.d, budget, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: GPR Exact, cov.RBF, N=%d D=%d fp64 (K build + Cholesky + solves + nlZ)"
                       % (args.n, args.d), "engine": "oracle port of pyGPs numpy/scipy path, host CPU",
                       "hyp": "changed every step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)
    return 0


def _emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


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
    eng.set_profile(True)      # two CUDA events around each step's trailing-update launches (same stream)
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

    # ---- roofline of the dominant kernel (rank 0; profiled evaluations outside the timed region) ----
    roof = None
    extra = {}
    if ctx.rank == 0:
        eng.set_profile(False)
        reps = steps               # measured live, inside the timed region above
        T = (N + NB - 1) // NB
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
            # int8 tensor ops against the int8 dense peak (2x the measured bf16 dense figure: same pipe, 32 instead of
            # 16 K-elements per instruction), cross-checked by the repo's own issue-rate probe.
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
            # DRAM traffic of the dominant launch (largest oz_syrk launch of an evaluation) from the committed ncu --set full
            # capture, next to the algorithmic bytes of that same launch (C lower triangle read + written once, slices once)
            traffic, traffic_note = None, None
            try:
                with open(os.path.join(ROOT, "profiles", "r1_oz_syrk_traffic.json")) as f:
                    tj = json.load(f)
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_note = {"launch": tj["launch"], "algorithmic_bytes": tj["algorithmic_bytes"]["total"],
                                "ratio": traffic / float(tj["algorithmic_bytes"]["total"]), "source": tj["source"]}
            except Exception:
                pass
            roof = {"bound": "tensor", "kernel": "oz_syrk_kernel<7,8,1,0> (trailing SYRK update: tcgen05.mma.kind::i8, TMEM accumulators)",
                    "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_note": traffic_note,
                    "peak_source": src,
                    "int8_probe": {"clk_per_mma_128x256x32": probe_clk, "tops": probe_tops},
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

    # ---- CPU baseline beside it (rank 0, N=1 only) ------------------------------------------
    cpu = None
    if ctx.rank == 0 and n_gpus == 1 and not args.no_cpu:
        v, _, sample, cores = cpu_port_measure(N, D, 40.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
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
                "config": {"workload": "C2: GPR Exact, cov.RBF, N=%d D=%d fp64 (K build + Cholesky + solves + nlZ)" % (N, D),
                           "parallelism": "replicas x%d (independent hyper-parameter vectors per GPU, no collective)" % n_gpus,
                           "l2": "working set (%.1f GiB factor) exceeds the 126 MB L2; no flush needed" % (8.0 * N * N / 2 ** 30),
                           "hyp": "changed every step"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                        "api": "pygps_b200.GPR().getPosterior(x, y, der=False) with pinned host arrays"},
                "gpu_launches": launches_total, "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
                "parity": {"nlZ": nlz_first, "reference_nlZ": ref_nlz, "rel_err": parity}}
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-der", action="store_true", help="skip the derivative-rate side measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
