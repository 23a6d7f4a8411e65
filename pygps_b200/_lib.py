"""ctypes binding of libgpk.so (include/gpk.h) - the only door to the GPU.

There is deliberately no CPU fallback here: if the shared library is missing or
no CUDA device is usable, every entry point raises.  The oracle under `oracle/`
is test infrastructure and is never imported from this package.
"""
import ctypes
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libgpk.so")

COV_RBF, COV_RBFARD, COV_MATERN = 0, 1, 2
# covariance-program ops (include/gpk.h: GPK_OP_*)
(OP_RBF, OP_RBFARD, OP_MATERN, OP_RBFUNIT, OP_RQ, OP_RQARD, OP_PERIODIC, OP_PIECEPOLY, OP_GABOR, OP_NOISE, OP_CONST,
 OP_LINEAR, OP_POLY, OP_PRE) = range(14)
OP_SUM, OP_PROD, OP_SCALE = 32, 33, 34
MODE_TRAIN, MODE_CROSS, MODE_SELF_TEST = 0, 1, 2
_MODES = {"train": MODE_TRAIN, "cross": MODE_CROSS, "self_test": MODE_SELF_TEST}

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)


class GpkStats(ctypes.Structure):
    _fields_ = [("total_ms", ctypes.c_double), ("kbuild_ms", ctypes.c_double),
                ("potrf_ms", ctypes.c_double), ("solve_ms", ctypes.c_double),
                ("deriv_ms", ctypes.c_double), ("syrk_ms", ctypes.c_double),
                ("syrk_flops", ctypes.c_double), ("launches", ctypes.c_int64),
                ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GpkError(RuntimeError):
    pass


class CovNode(ctypes.Structure):
    """gpk_cov_node (include/gpk.h)"""
    _fields_ = [("op", ctypes.c_int32), ("a", ctypes.c_int32), ("b", ctypes.c_int32), ("hyp0", ctypes.c_int32),
                ("para", ctypes.c_double)]


def _node_array(nodes):
    arr = (CovNode * len(nodes))()
    for i, (op, a, b, h0, para) in enumerate(nodes):
        arr[i].op, arr[i].a, arr[i].b, arr[i].hyp0, arr[i].para = int(op), int(a), int(b), int(h0), float(para)
    return arr


_lib = None
_lib_lock = threading.Lock()

# every symbol include/gpk.h declares: (name, restype, argtypes)
_H = ctypes.c_void_p
_I, _L, _D = ctypes.c_int, ctypes.c_int64, ctypes.c_double
PROTOTYPES = [
    ("gpk_version", _I, []),
    ("gpk_strerror", ctypes.c_char_p, [_I]),
    ("gpk_device_count", _I, [c_int_p]),
    ("gpk_device_memory", _I, [_I, ctypes.POINTER(_L), ctypes.POINTER(_L)]),
    ("gpk_create", _I, [_I, ctypes.POINTER(_H)]),
    ("gpk_destroy", _I, [_H]),
    ("gpk_last_error", ctypes.c_char_p, [_H]),
    ("gpk_last_stats", _I, [_H, ctypes.POINTER(GpkStats)]),
    ("gpk_set_profile", _I, [_H, _I]),
    ("gpk_cov_matrix", _I, [_H, _I, _I, c_double_p, _I, c_double_p, _L, c_double_p, _L, _I, _I, _I, c_double_p]),
    ("gpk_cov_matrix_prog", _I, [_H, ctypes.POINTER(CovNode), _I, c_double_p, _I, c_double_p, _L, c_double_p, _L, _I, _I, _I,
                                 c_double_p]),
    ("gpk_set_pre", _I, [_H, c_double_p, _L]),
    ("gpk_exact_eval_prog", _I, [_H, ctypes.POINTER(CovNode), _I, c_double_p, _I, _D, c_double_p, _I,
                                 c_double_p, c_double_p, c_double_p, c_double_p]),
    ("gpk_potrf", _I, [_H, c_double_p, _L, c_double_p, c_double_p]),
    ("gpk_set_factor", _I, [_H, c_double_p, _L, c_double_p]),
    ("gpk_potrs", _I, [_H, c_double_p, _L, _L, c_double_p]),
    ("gpk_set_data", _I, [_H, c_double_p, _L, _I]),
    ("gpk_exact_eval", _I, [_H, _I, _I, c_double_p, _I, _D, c_double_p, _I,
                            c_double_p, c_double_p, c_double_p, c_double_p]),
    ("gpk_get_factor", _I, [_H, c_double_p]),
    ("gpk_predict", _I, [_H, c_double_p, _L, c_double_p, c_double_p]),
    ("gpk_fitc_eval", _I, [_H, _I, _I, c_double_p, _I, _D, c_double_p, _L, c_double_p, _I,
                           c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    ("gpk_fitc_predict", _I, [_H, c_double_p, _L, c_double_p, c_double_p]),
    ("gpk_ep_eval", _I, [_H, _I, _I, c_double_p, _I, c_double_p, c_double_p, c_double_p, c_double_p, _I, _I,
                         c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int_p]),
    ("gpk_dist_unique_id", _I, [ctypes.c_char_p, ctypes.c_char_p]),
    ("gpk_dist_init", _I, [_H, ctypes.c_char_p, _I, _I, ctypes.c_char_p]),
    ("gpk_dist_finalize", _I, [_H]),
    ("gpk_dist_reserve", _I, [_H, _I]),
    ("gpk_exact_eval_dist", _I, [_H, _I, _I, c_double_p, _I, _D, c_double_p, c_double_p, c_double_p]),
    ("gpk_dist_gather_factor", _I, [_H]),
    ("gpk_exact_eval_dist_der", _I, [_H, _I, _I, c_double_p, _I, _D, c_double_p, c_double_p, c_double_p, c_double_p,
                                     c_double_p]),
    ("gpk_bench_dmma", _I, [_H, _I, _I, _I, c_double_p, c_double_p]),
    ("gpk_bench_syrk", _I, [_H, _L, _I, _I, c_double_p, c_double_p]),
    ("gpk_bench_copy", _I, [_H, _L, _I, c_double_p]),
    ("gpk_dbg_gemm_nt", _I, [_H, _I, _L, _L, _L, c_double_p, c_double_p, c_double_p]),
    ("gpk_dbg_diag", _I, [_H, c_double_p, c_double_p, c_double_p, c_double_p, c_int_p]),
    ("gpk_dbg_i8_tile", _I, [_H, _I, _I, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _I]),
    ("gpk_bench_i8_rate", _I, [_H, _I, _I, _I, _I, _I, _I, _I, c_double_p, c_double_p]),
    ("gpk_dbg_oz_syrk", _I, [_H, _L, _I, c_double_p, c_double_p, _I, _I, c_double_p]),
]


def lib_path():
    return _LIBPATH


def load():
    """Load libgpk.so and attach prototypes.  Raises GpkError when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIBPATH):
            raise GpkError(
                "libgpk.so is not built (%s). Run `python -m pygps_b200.build`; "
                "this package has no CPU fallback." % _LIBPATH)
        lib = ctypes.CDLL(_LIBPATH)
        for name, res, args in PROTOTYPES:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def as_f64(a, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise Exception("%s must be a 2-d array" % name)
    return a


def _no_engine():
    return None


class Engine(object):
    """One GPU handle.  Not re-entrant; one in-flight call at a time.

    A handle cannot be copied or sent to another process: copy.deepcopy / pickle of an object that holds one
    (a model, its inference method) yields None in its place, and the copy creates its own handle on first use
    - reference models are plain Python objects and survive deepcopy / joblib / multiprocessing, so must these."""

    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_no_engine, ())

    def __init__(self, device=None):
        self._lib = load()
        if device is None:
            device = int(os.environ.get("PYGPS_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            cnt = ctypes.c_int(0)
            self._lib.gpk_device_count(ctypes.byref(cnt))
            if cnt.value > 0:
                device %= cnt.value
        self.device = device
        h = _H()
        rc = self._lib.gpk_create(device, ctypes.byref(h))
        if rc != 0:
            raise GpkError("gpk_create(device=%d) failed: %s - a CUDA device is required, there is no CPU fallback"
                           % (device, self._lib.gpk_strerror(rc).decode()))
        self._h = h
        self.epoch = 0              # bumped whenever the resident factor changes
        self._finalizer = weakref.finalize(self, self._lib.gpk_destroy, h)

    # -- helpers ------------------------------------------------------------
    def _check(self, rc, what):
        if rc == 0:
            return
        if rc > 0:
            # LAPACK-style info: same exception type and text as Core/tools.py:77
            raise np.linalg.LinAlgError("kernel matrix not positive definite, even with jitter.")
        msg = self._lib.gpk_strerror(rc).decode()
        if rc == -2:
            msg += ": " + self._lib.gpk_last_error(self._h).decode()
        raise GpkError("%s failed (%d): %s" % (what, rc, msg))

    def stats(self):
        s = GpkStats()
        self._lib.gpk_last_stats(self._h, ctypes.byref(s))
        return s.as_dict()

    def set_profile(self, on):
        self._lib.gpk_set_profile(self._h, int(bool(on)))

    def _retire_factor(self):
        """Called before the resident factor is overwritten.  Posteriors that still point at it
        notice through the epoch and rebuild their factor on demand (inf.postStruct._materialize)."""
        self.epoch += 1

    # -- covariance -----------------------------------------------------------
    def cov_matrix(self, kind, matern_d, hyp, x, z, mode, der=-1):
        hyp = np.ascontiguousarray(hyp, dtype=np.float64)
        m = _MODES[mode]
        if m == MODE_SELF_TEST:
            z = as_f64(z, "z")
            out = np.empty((z.shape[0], 1))
            rc = self._lib.gpk_cov_matrix(self._h, kind, matern_d, _dp(hyp), hyp.size, None, 0, _dp(z),
                                          z.shape[0], z.shape[1], m, der, _dp(out))
        elif m == MODE_TRAIN:
            x = as_f64(x, "x")
            out = np.empty((x.shape[0], x.shape[0]))
            rc = self._lib.gpk_cov_matrix(self._h, kind, matern_d, _dp(hyp), hyp.size, _dp(x), x.shape[0],
                                          None, 0, x.shape[1], m, der, _dp(out))
        else:
            x = as_f64(x, "x")
            z = as_f64(z, "z")
            if x.shape[1] != z.shape[1]:
                raise Exception("x and z must have the same number of columns")
            out = np.empty((x.shape[0], z.shape[0]))
            rc = self._lib.gpk_cov_matrix(self._h, kind, matern_d, _dp(hyp), hyp.size, _dp(x), x.shape[0],
                                          _dp(z), z.shape[0], x.shape[1], m, der, _dp(out))
        self._check(rc, "gpk_cov_matrix")
        return out

    def cov_matrix_prog(self, nodes, hyp, x, z, mode, der=-1):
        """getCovMatrix / getDerMatrix of a covariance program (composite kernels, csrc/covprog.cu)."""
        hyp = np.ascontiguousarray(hyp, dtype=np.float64)
        arr = _node_array(nodes)
        m = _MODES[mode]
        x = None if (x is None or m == MODE_SELF_TEST) else as_f64(x, "x")
        z = None if (z is None or m == MODE_TRAIN) else as_f64(z, "z")
        if m == MODE_CROSS and x.shape[1] != z.shape[1]:
            raise Exception("x and z must have the same number of columns")
        D = (x if x is not None else z).shape[1]
        n = 0 if x is None else x.shape[0]
        mm = 0 if z is None else z.shape[0]
        out = np.empty((mm, 1)) if m == MODE_SELF_TEST else np.empty((n, n if m == MODE_TRAIN else mm))
        rc = self._lib.gpk_cov_matrix_prog(self._h, arr, len(nodes), _dp(hyp), hyp.size, _dp(x), n, _dp(z), mm, D, m,
                                           int(der), _dp(out))
        self._check(rc, "gpk_cov_matrix_prog")
        return out

    def set_pre(self, K):
        K = as_f64(K, "precomputed training matrix")
        if K.shape[0] != K.shape[1]:
            raise Exception("precomputed training matrix must be square")
        self._check(self._lib.gpk_set_pre(self._h, _dp(K), K.shape[0]), "gpk_set_pre")

    def exact_eval_prog(self, nodes, hyp, log_sn, ymm, want_der):
        hyp = np.ascontiguousarray(hyp, dtype=np.float64)
        ymm = np.ascontiguousarray(ymm, dtype=np.float64).reshape(-1)
        arr = _node_array(nodes)
        self._retire_factor()
        alpha = np.empty((ymm.size, 1))
        nlZ = ctypes.c_double(0.0)
        dcov = np.zeros(max(hyp.size, 1))
        dlik = np.zeros(1)
        rc = self._lib.gpk_exact_eval_prog(self._h, arr, len(nodes), _dp(hyp), hyp.size, float(log_sn), _dp(ymm),
                                           1 if want_der else 0, ctypes.byref(nlZ), _dp(alpha), _dp(dcov), _dp(dlik))
        self._check(rc, "gpk_exact_eval_prog")
        return np.float64(nlZ.value), alpha, dcov[:hyp.size], dlik

    # -- jitchol / solve_chol -------------------------------------------------
    def potrf(self, A, want_factor=True):
        A = as_f64(A, "A")
        n = A.shape[0]
        self._retire_factor()
        R = np.empty((n, n)) if want_factor else None
        ld = ctypes.c_double(0.0)
        rc = self._lib.gpk_potrf(self._h, _dp(A), n, _dp(R), ctypes.byref(ld))
        if rc > 0:
            if np.any(np.diag(A) <= 0.0):
                raise np.linalg.LinAlgError(
                    "kernel matrix not positive definite: non-positive diagonal elements")
        self._check(rc, "gpk_potrf")
        return R, ld.value

    def set_factor(self, R):
        """Make the upper factor R (A = R'R) the resident one (solve_chol with a factor that is not on the GPU)."""
        R = as_f64(R, "R")
        if R.shape[0] != R.shape[1]:
            raise Exception("factor must be square")
        self._retire_factor()
        ld = ctypes.c_double(0.0)
        rc = self._lib.gpk_set_factor(self._h, _dp(R), R.shape[0], ctypes.byref(ld))
        if rc > 0:
            raise np.linalg.LinAlgError("not a Cholesky factor: non-positive diagonal element")
        self._check(rc, "gpk_set_factor")
        return ld.value

    def potrs(self, B):
        B = as_f64(B, "B")
        X = np.empty_like(B)
        rc = self._lib.gpk_potrs(self._h, _dp(B), B.shape[0], B.shape[1], _dp(X))
        if rc == -1:
            raise Exception("Wrong sizes of matrix arguments in solve_chol.py")
        self._check(rc, "gpk_potrs")
        return X

    # -- exact inference ------------------------------------------------------
    def set_data(self, x):
        x = as_f64(x, "x")
        self._retire_factor()
        rc = self._lib.gpk_set_data(self._h, _dp(x), x.shape[0], x.shape[1])
        self._check(rc, "gpk_set_data")
        self.n, self.D = x.shape

    def exact_eval(self, kind, matern_d, hyp, log_sn, ymm, want_der):
        hyp = np.ascontiguousarray(hyp, dtype=np.float64)
        ymm = np.ascontiguousarray(ymm, dtype=np.float64).reshape(-1)
        n = ymm.size
        self._retire_factor()
        alpha = np.empty((n, 1))
        nlZ = ctypes.c_double(0.0)
        dcov = np.zeros(max(hyp.size, 1))
        dlik = np.zeros(1)
        rc = self._lib.gpk_exact_eval(self._h, kind, matern_d, _dp(hyp), hyp.size, float(log_sn), _dp(ymm),
                                      1 if want_der else 0, ctypes.byref(nlZ), _dp(alpha), _dp(dcov), _dp(dlik))
        self._check(rc, "gpk_exact_eval")
        return np.float64(nlZ.value), alpha, dcov[:hyp.size], dlik

    def get_factor(self, n):
        R = np.empty((n, n))
        rc = self._lib.gpk_get_factor(self._h, _dp(R))
        self._check(rc, "gpk_get_factor")
        return R

    def predict(self, xs):
        xs = as_f64(xs, "xs")
        ns = xs.shape[0]
        ka = np.empty((ns, 1))
        fs2 = np.empty((ns, 1))
        rc = self._lib.gpk_predict(self._h, _dp(xs), ns, _dp(ka), _dp(fs2))
        self._check(rc, "gpk_predict")
        return ka, fs2

    # -- EP classification --------------------------------------------------------------
    def ep_eval(self, kind, matern_d, hyp, mvec, y, ttau, tnu, use_last, want_der):
        hyp = np.ascontiguousarray(hyp, dtype=np.float64)
        mvec = np.ascontiguousarray(mvec, dtype=np.float64).reshape(-1)
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        n = y.size
        ttau = np.ascontiguousarray(ttau, dtype=np.float64).reshape(-1).copy() if use_last else np.zeros(n)
        tnu = np.ascontiguousarray(tnu, dtype=np.float64).reshape(-1).copy() if use_last else np.zeros(n)
        self._retire_factor()
        alpha = np.empty((n, 1)); sW = np.empty((n, 1)); dlz = np.zeros((n, 1))
        nlZ = ctypes.c_double(0.0)
        dcov = np.zeros(max(hyp.size, 1))
        sweeps = ctypes.c_int(0)
        rc = self._lib.gpk_ep_eval(self._h, kind, matern_d, _dp(hyp), hyp.size, _dp(mvec), _dp(y), _dp(ttau), _dp(tnu),
                                   1 if use_last else 0, 1 if want_der else 0, ctypes.byref(nlZ), _dp(alpha), _dp(sW),
                                   _dp(dcov), _dp(dlz), ctypes.byref(sweeps))
        self._check(rc, "gpk_ep_eval")
        return (np.float64(nlZ.value), alpha, sW, dcov[:hyp.size], dlz, ttau.reshape(-1, 1), tnu.reshape(-1, 1),
                sweeps.value)

    # -- one evaluation sharded over several GPUs (one process per GPU) ------------------
    @staticmethod
    def nccl_path():
        """libnccl.so.2 as shipped with PyTorch (resolved without importing torch's CUDA runtime)."""
        try:
            import nvidia.nccl
            base = os.path.dirname(nvidia.nccl.__file__) if getattr(nvidia.nccl, "__file__", None) else \
                list(nvidia.nccl.__path__)[0]
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
        except Exception:
            pass
        return "libnccl.so.2"

    def dist_unique_id(self):
        buf = ctypes.create_string_buffer(128)
        rc = self._lib.gpk_dist_unique_id(self.nccl_path().encode(), buf)
        self._check(rc, "gpk_dist_unique_id")
        return buf.raw

    def dist_init(self, rank, world, uid=None):
        rc = self._lib.gpk_dist_init(self._h, self.nccl_path().encode(), rank, world, uid)
        self._check(rc, "gpk_dist_init")
        self.rank, self.world = rank, world

    def exact_eval_dist(self, kind, matern_d, hyp, log_sn, ymm):
        hyp = np.ascontiguousarray(hyp, dtype=np.float64)
        ymm = np.ascontiguousarray(ymm, dtype=np.float64).reshape(-1)
        self._retire_factor()
        alpha = np.empty((ymm.size, 1))
        nlZ = ctypes.c_double(0.0)
        rc = self._lib.gpk_exact_eval_dist(self._h, kind, matern_d, _dp(hyp), hyp.size, float(log_sn), _dp(ymm),
                                           ctypes.byref(nlZ), _dp(alpha))
        self._check(rc, "gpk_exact_eval_dist")
        return np.float64(nlZ.value), alpha

    def exact_eval_dist_der(self, kind, matern_d, hyp, log_sn, ymm):
        hyp = np.ascontiguousarray(hyp, dtype=np.float64)
        ymm = np.ascontiguousarray(ymm, dtype=np.float64).reshape(-1)
        self._retire_factor()
        alpha = np.empty((ymm.size, 1))
        nlZ = ctypes.c_double(0.0)
        dcov = np.zeros(max(hyp.size, 1))
        dlik = np.zeros(1)
        rc = self._lib.gpk_exact_eval_dist_der(self._h, kind, matern_d, _dp(hyp), hyp.size, float(log_sn), _dp(ymm),
                                               ctypes.byref(nlZ), _dp(alpha), _dp(dcov), _dp(dlik))
        self._check(rc, "gpk_exact_eval_dist_der")
        return np.float64(nlZ.value), alpha, dcov[:hyp.size], dlik

    def dist_reserve(self, level):
        self._check(self._lib.gpk_dist_reserve(self._h, int(level)), "gpk_dist_reserve")

    def dist_gather_factor(self):
        self._check(self._lib.gpk_dist_gather_factor(self._h), "gpk_dist_gather_factor")

    # -- FITC -----------------------------------------------------------------
    def fitc_eval(self, kind, matern_d, hyp, log_sn, u, ymm, want_der):
        hyp = np.ascontiguousarray(hyp, dtype=np.float64)
        u = as_f64(u, "inducing inputs")
        ymm = np.ascontiguousarray(ymm, dtype=np.float64).reshape(-1)
        M, n = u.shape[0], ymm.size
        self._retire_factor()
        alpha = np.empty((M, 1))
        Lp = np.empty((M, M))
        nlZ = ctypes.c_double(0.0)
        dcov = np.zeros(max(hyp.size, 1))
        dlik = np.zeros(1)
        al = np.zeros((n, 1))
        rc = self._lib.gpk_fitc_eval(self._h, kind, matern_d, _dp(hyp), hyp.size, float(log_sn), _dp(u), M,
                                     _dp(ymm), 1 if want_der else 0, ctypes.byref(nlZ), _dp(alpha), _dp(Lp),
                                     _dp(dcov), _dp(dlik), _dp(al))
        self._check(rc, "gpk_fitc_eval")
        return np.float64(nlZ.value), alpha, Lp, dcov[:hyp.size], dlik, al

    def fitc_predict(self, xs):
        xs = as_f64(xs, "xs")
        ns = xs.shape[0]
        ka = np.empty((ns, 1))
        fs2 = np.empty((ns, 1))
        rc = self._lib.gpk_fitc_predict(self._h, _dp(xs), ns, _dp(ka), _dp(fs2))
        self._check(rc, "gpk_fitc_predict")
        return ka, fs2

    # -- measurement ----------------------------------------------------------
    def bench_dmma(self, shape, warps=8, iters=20000):
        tf, ms = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self._lib.gpk_bench_dmma(self._h, shape, warps, iters, ctypes.byref(tf), ctypes.byref(ms)),
                    "gpk_bench_dmma")
        return tf.value, ms.value

    def bench_syrk(self, n, k=128, reps=5):
        tf, ms = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self._lib.gpk_bench_syrk(self._h, n, k, reps, ctypes.byref(ms), ctypes.byref(tf)),
                    "gpk_bench_syrk")
        return ms.value, tf.value

    def bench_copy(self, nbytes=1 << 30, reps=5):
        g = ctypes.c_double(0)
        self._check(self._lib.gpk_bench_copy(self._h, nbytes, reps, ctypes.byref(g)), "gpk_bench_copy")
        return g.value

    def dbg_gemm_nt(self, mode, A, B, C):
        """Column-major (Fortran-ordered) A (M,K), B (N,K), C (M,N); returns the new C (mode 7, the head pair of the
        panel chain: the new C and the overwritten A)."""
        A = np.array(A, dtype=np.float64, order="F", copy=True)
        B = np.asfortranarray(B, dtype=np.float64)
        C = np.array(C, dtype=np.float64, order="F", copy=True)
        M, K = A.shape
        N = B.shape[0]
        rc = self._lib.gpk_dbg_gemm_nt(self._h, mode, M, N, K, A.ctypes.data_as(c_double_p),
                                       B.ctypes.data_as(c_double_p), C.ctypes.data_as(c_double_p))
        self._check(rc, "gpk_dbg_gemm_nt")
        return (C, A) if mode == 7 else C

    def dbg_i8_tile(self, A, B, a_tmem=0):
        """int32 C = A (128,K) @ B (N,K)^T through one tcgen05.mma.kind::i8 tile; A, B int8 or uint8 arrays.
        a_tmem 1: A staged through TMEM by tcgen05.cp (one region per k-step), 2: one reused region."""
        fmt = (1 if A.dtype == np.uint8 else 0) | (2 if B.dtype == np.uint8 else 0) | (4 if a_tmem == 1 else 0) | (8 if a_tmem == 2 else 0)
        A = np.ascontiguousarray(A)
        B = np.ascontiguousarray(B)
        if A.dtype.itemsize != 1 or B.dtype.itemsize != 1:
            raise TypeError("int8 / uint8 operands expected")
        C = np.zeros((128, B.shape[0]), dtype=np.int32)
        rc = self._lib.gpk_dbg_i8_tile(self._h, B.shape[0], A.shape[1], A.ctypes.data, B.ctypes.data, C.ctypes.data, fmt)
        self._check(rc, "gpk_dbg_i8_tile")
        return C

    def bench_i8_rate(self, N, iters, lbo, sbo, astep=0, same_acc=0, ctas=148):
        """(clocks per MMA, int8 TOP/s) of back-to-back tcgen05.mma.kind::i8 128xNx32 from `ctas` CTAs."""
        v = ctypes.c_double(0)
        t = ctypes.c_double(0)
        rc = self._lib.gpk_bench_i8_rate(self._h, N, iters, lbo, sbo, astep, same_acc, ctas, ctypes.byref(v), ctypes.byref(t))
        self._check(rc, "gpk_bench_i8_rate")
        return v.value, t.value

    def dbg_oz_syrk(self, P, C, mode=0, reps=1):
        """lower(C) - P P' through the int8 tensor-core path (mode 0) or the DMMA path (mode 1); returns (C_new, ms)."""
        P = np.asfortranarray(P, dtype=np.float64)
        C = np.array(C, dtype=np.float64, order="F", copy=True)
        ms = ctypes.c_double(0)
        rc = self._lib.gpk_dbg_oz_syrk(self._h, P.shape[0], P.shape[1], P.ctypes.data_as(c_double_p),
                                       C.ctypes.data_as(c_double_p), mode, reps, ctypes.byref(ms))
        self._check(rc, "gpk_dbg_oz_syrk")
        return C, ms.value

    def dbg_diag(self, A):
        A = np.asfortranarray(A, dtype=np.float64)
        L = np.zeros((128, 128), order="F")
        Li = np.zeros((128, 128), order="F")
        ld = ctypes.c_double(0)
        info = ctypes.c_int(0)
        rc = self._lib.gpk_dbg_diag(self._h, A.ctypes.data_as(c_double_p), L.ctypes.data_as(c_double_p),
                                    Li.ctypes.data_as(c_double_p), ctypes.byref(ld), ctypes.byref(info))
        self._check(rc, "gpk_dbg_diag")
        return L, Li, ld.value, info.value


# ---------------------------------------------------------------------------------------------
# devices of this process
_devices = None


def device_count():
    c = ctypes.c_int(0)
    load().gpk_device_count(ctypes.byref(c))
    return c.value


def set_devices(devices):
    """Restrict (or order) the GPUs this process uses for multi-GPU work: random restarts of the optimizers and
    sharded evaluations.  None = every visible device."""
    global _devices
    _devices = None if devices is None else [int(d) for d in devices]


def visible_devices():
    """CUDA ordinals this process may use.  Under a one-process-per-GPU launcher (torchrun: WORLD_SIZE > 1) a process
    owns exactly its LOCAL_RANK device; otherwise every visible device, or the list given to set_devices()."""
    if _devices is not None:
        return list(_devices)
    n = device_count()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and "LOCAL_RANK" in os.environ:
        return [int(os.environ["LOCAL_RANK"]) % max(n, 1)]
    return list(range(n))


def device_memory(device):
    """(free, total) bytes of HBM on `device`."""
    f, t = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = load().gpk_device_memory(int(device), ctypes.byref(f), ctypes.byref(t))
    if rc != 0:
        raise GpkError("gpk_device_memory(%d) failed" % device)
    return f.value, t.value


class ShardedEngine(object):
    """ONE evaluation sharded over several GPUs of this process (BASELINE config 3): one Engine and one thread per
    GPU, bound together by an NCCL communicator (gpk_dist_init from every thread at once; ctypes releases the GIL, so
    the ranks really run concurrently).  The factor is distributed by block-cyclic block columns; every rank
    broadcasts its solved panels over NVLink (csrc/dist.cu)."""

    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_no_engine, ())

    def __init__(self, devices):
        from concurrent.futures import ThreadPoolExecutor
        self.devices = [int(d) for d in devices]
        if len(set(self.devices)) != len(self.devices) or not self.devices:
            raise GpkError("ShardedEngine needs distinct devices")
        self.world = len(self.devices)
        self.engines = [Engine(d) for d in self.devices]
        self._pool = ThreadPoolExecutor(self.world)
        uid = self.engines[0].dist_unique_id() if self.world > 1 else None
        self._all(lambda r, e: e.dist_init(r, self.world, uid))
        self.epoch = 0

    def _all(self, fn):
        futs = [self._pool.submit(fn, r, e) for r, e in enumerate(self.engines)]
        return [f.result() for f in futs]

    def set_data(self, x):
        x = as_f64(x, "x")
        self._all(lambda r, e: e.set_data(x))
        self.n, self.D = x.shape

    def exact_eval(self, kind, matern_d, hyp, log_sn, ymm, want_der=False):
        """nlZ, alpha[, dcov, dlik] - same values on every rank; rank 0's are returned."""
        self.epoch += 1
        # allocate on every rank FIRST and join (gpk_dist_reserve): no device allocation inside the collective phase
        self._all(lambda r, e: e.dist_reserve(2 if want_der else 0))
        if want_der:
            out = self._all(lambda r, e: e.exact_eval_dist_der(kind, matern_d, hyp, log_sn, ymm))
            return out[0]
        out = self._all(lambda r, e: e.exact_eval_dist(kind, matern_d, hyp, log_sn, ymm))
        return out[0][0], out[0][1], np.zeros(len(hyp)), np.zeros(1)

    def get_factor(self, n):
        self._all(lambda r, e: e.dist_reserve(1))
        self._all(lambda r, e: e.dist_gather_factor())
        return self.engines[0].get_factor(n)

    def predict(self, xs):
        """Test points are split over the ranks; every rank solves its share against its replica of the factor
        (gathered once per posterior, gpk_dist_gather_factor)."""
        xs = as_f64(xs, "xs")
        ns = xs.shape[0]
        cuts = [(ns * r) // self.world for r in range(self.world + 1)]

        self._all(lambda r, e: e.dist_reserve(1))
        self._all(lambda r, e: e.dist_gather_factor())

        def part(r, e):
            if cuts[r + 1] == cuts[r]:
                return np.empty((0, 1)), np.empty((0, 1))
            return e.predict(xs[cuts[r]:cuts[r + 1]])
        out = self._all(part)
        return np.vstack([o[0] for o in out]), np.vstack([o[1] for o in out])

    def stats(self):
        return [e.stats() for e in self.engines]

    def close(self):
        self._pool.shutdown(wait=True)


_shared = {}
_shared_lock = threading.Lock()


def shared_engine(device=None):
    """Process-wide engine used by the stateless entry points (getCovMatrix, jitchol, solve_chol)."""
    key = device
    with _shared_lock:
        e = _shared.get(key)
        if e is None:
            e = Engine(device)
            _shared[key] = e
        return e
