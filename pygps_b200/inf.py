"""Inference engines on the accelerated path: Exact and FITC_Exact.

Mirror of pyGPs.Core.inf (/root/reference/pyGPs/Core/inf.py): same
`evaluate(meanfunc, covfunc, likfunc, x, y, nargout)` contract (:140-172), same
result containers postStruct (:59-89) and dnlZStruct (:93-128), same exceptions.
One evaluation is ONE foreign call (gpk_exact_eval / gpk_fitc_eval): the kernel
matrix, its factor and the solves never leave the GPU; post.L is fetched lazily.
"""
import logging
from copy import copy

import numpy as np

from . import _lib, cov, lik
from .tools import jitchol, solve_chol

np.seterr(all='ignore')          # the reference does this at import (Core/inf.py:56)


_SCRATCH = []


def _scratch_engine():
    """A second GPU handle used only to rebuild a factor that was overwritten on the model's handle."""
    if not _SCRATCH:
        _SCRATCH.append(_lib.Engine())
    return _SCRATCH[0]


def _x_signature(x):
    """Cheap fingerprint of the training inputs (shape + a few entries + sum): detects in-place edits."""
    return (x.shape, float(x.flat[0]), float(x.flat[-1]), float(x.sum()))


def _vec_signature(post):
    """Fingerprint of the posterior vectors as the evaluation produced them: a caller-edited alpha / sW no longer
    matches, and prediction then uses the host arrays the way the reference does (Core/gp.py:402-419)."""
    a, w = np.asarray(post.alpha), np.asarray(post.sW)
    return (a.shape, w.shape, float(a.sum()), float(np.abs(a).sum()), float(w.sum()))


def _seal(post, x):
    post._x, post._xsig = x, _x_signature(x)
    post._vsig = _vec_signature(post)


class postStruct(object):
    """Posterior parameters alpha, sW, L (Core/inf.py:59-89).

    L is device-backed: it is copied out of GPU memory the first time it is read
    (an (n,n) upper-triangular C-ordered array, exactly what the reference stores)."""

    def __init__(self):
        self.alpha = np.array([])
        self.sW = np.array([])
        self._L = np.array([])
        self._engine = None      # engine that holds the factor this posterior describes
        self._epoch = -1
        self._n = 0
        self._spec = None        # what the resident factor was built from (for predict / rebuild)
        self._x = None           # training inputs the factor was built from
        self._xsig = None
        self._vsig = None        # fingerprint of alpha / sW as evaluated (see _vec_signature)

    # -- lazy factor ----------------------------------------------------------
    def _resident(self):
        return self._engine is not None and self._engine.epoch == self._epoch

    def _materialize(self):
        """post.L on first touch: copied out of the GPU if the factor is still resident, otherwise
        rebuilt there from (x, hyper-parameters) - the factorisation is deterministic, so the result is
        the same array the reference would have kept (at the price of one more evaluation)."""
        if self._L is None:
            if self._resident():
                self._L = self._engine.get_factor(self._n)
            else:
                if self._x is None or self._spec is None or _x_signature(self._x) != self._xsig:
                    raise RuntimeError("posterior factor is no longer resident on the GPU and the training "
                                       "inputs it was built from have changed")
                _, kind, md, hyp, log_sn = self._spec
                eng = _scratch_engine()
                eng.set_data(self._x)
                if kind == 'prog':
                    eng.exact_eval_prog(list(md), list(hyp), log_sn, np.zeros(self._x.shape[0]), False)
                else:
                    eng.exact_eval(kind, md, list(hyp), log_sn, np.zeros(self._x.shape[0]), False)
                self._L = eng.get_factor(self._n)
        return self._L

    def _getL(self):
        return self._materialize()

    def _setL(self, value):
        self._L = value
        self._engine = None
    L = property(_getL, _setL)

    def __deepcopy__(self, memo):
        # GP.getPosterior deep-copies the posterior every call (Core/gp.py:338,344); copying a
        # lazy handle instead of an N x N matrix keeps that cheap.
        other = postStruct()
        other.alpha = np.array(self.alpha, copy=True)
        other.sW = np.array(self.sW, copy=True)
        other._L = None if self._L is None else np.array(self._L, copy=True)
        other._engine, other._epoch, other._n, other._spec = self._engine, self._epoch, self._n, self._spec
        other._x, other._xsig, other._vsig = self._x, self._xsig, self._vsig
        return other

    def __getstate__(self):
        # pickling (joblib, multiprocessing): the factor is materialised, the GPU handle stays behind
        d = dict(self.__dict__)
        d["_L"] = self._materialize()
        d["_engine"], d["_epoch"] = None, -1
        return d

    def __repr__(self):
        return ("posterior: to get the parameters of the posterior distribution use:\n"
                "model.posterior.alpha\nmodel.posterior.L\nmodel.posterior.sW\n"
                "See documentation and gpml book chapter 2.3 and chapter 3.4.3 for these parameters.")

    def __str__(self):
        return ("posterior distribution described by alpha, sW and L\n"
                "See documentation and gpml book chapter 2.3 and chapter 3.4.3 for these parameters\n"
                "alpha:\n" + str(self.alpha) + "\nL:\n" + str(self.L) + "\nsW:\n" + str(self.sW))


class dnlZStruct(object):
    """Derivatives of nlZ w.r.t. mean / cov / lik hyper-parameters (Core/inf.py:93-128)."""

    def __init__(self, m, c, l):
        self.mean = []
        self.cov = []
        self.lik = []
        if m.hyp is not None:
            self.mean = [0 for i in range(len(m.hyp))]
        if c.hyp is not None:
            self.cov = [0 for i in range(len(c.hyp))]
        if l.hyp is not None:
            self.lik = [0 for i in range(len(l.hyp))]

    def __str__(self):
        return ("Derivatives of mean, cov and lik functions:\nmean:" + str(self.mean) + "\ncov:"
                + str(self.cov) + "\nlik:" + str(self.lik))

    def __repr__(self):
        return ("dnlZ: to get the derivatives of mean, cov and lik functions use:\n"
                "model.dnlZ.mean\nmodel.dnlZ.cov\nmodel.dnlZ.lik")

    def accumulateDnlZ(self, dnlZObject):
        self.mean = [a + b for a, b in zip(self.mean, dnlZObject.mean)]
        self.cov = [a + b for a, b in zip(self.cov, dnlZObject.cov)]
        self.lik = [a + b for a, b in zip(self.lik, dnlZObject.lik)]
        return self


class Inference(object):
    """Base class (Core/inf.py:133-172)."""

    def __init__(self):
        self.logger = logging.getLogger(__name__)
        self._engine = None

    devices = None        # CUDA ordinals this inference method may use (None: the process default)
    shard = None          # Exact only - None: shard an evaluation over `devices` when its factor does not fit one GPU;
                          # True / False: always / never

    def _device_list(self):
        return list(self.devices) if self.devices else _lib.visible_devices()

    def _get_engine(self):
        if getattr(self, '_engine', None) is None:
            self._engine = _lib.Engine(self.devices[0] if self.devices else None)
        return self._engine

    def evaluate(self, meanfunc, covfunc, likfunc, x, y, nargout=1):
        pass


def _mean_derivs(meanfunc, x, vec, dnlZ):
    """dnlZ.mean[i] = -dm_i' vec (Core/inf.py:378-381, :449-451): O(n) host products."""
    for i in range(len(meanfunc.hyp)):
        dnlZ.mean[i] = np.float64(np.dot(-meanfunc.getDerMatrix(x, i).T, vec)[0, 0])


class Exact(Inference):
    """Exact inference for a GP with Gaussian likelihood (Core/inf.py:345-384)."""

    def __init__(self, devices=None, shard=None):
        self.name = "Exact inference"
        self._engine = None
        self._sharded = None
        self.devices = devices
        self.shard = shard

    def _wants_sharding(self, n):
        """Route the evaluation to the sharded path (gpk_exact_eval_dist) when asked to, or when the N x N factor
        (plus the int8 slices of one block) does not fit one GPU of the device list."""
        devs = self._device_list()
        if self.shard:
            return True                      # (one device: the sharded code path with world size 1, no NCCL)
        if self.shard is False or len(devs) < 2:
            return False
        npad = -(-n // 128) * 128
        need = 8 * npad * npad + 8 * npad * 1152 + (1 << 30)
        try:
            free, total = _lib.device_memory(devs[0])
        except Exception:
            return False
        return need > 0.85 * total

    def _get_sharded(self):
        devs = self._device_list()
        if getattr(self, '_sharded', None) is None or self._sharded.devices != devs:
            self._sharded = _lib.ShardedEngine(devs)
        return self._sharded

    def evaluate(self, meanfunc, covfunc, likfunc, x, y, nargout=1):
        if not isinstance(likfunc, lik.Gauss):
            raise Exception('Exact inference only possible with Gaussian likelihood')
        n, D = x.shape
        m = meanfunc.getMean(x)
        sn2 = np.exp(2 * likfunc.hyp[0])
        if isinstance(covfunc, cov.FITCOfKernel):
            raise Exception('inf.Exact needs a plain covariance function (use inf.FITC_Exact with cov.FITCOfKernel)')
        spec = covfunc._device_spec()
        if spec is None:
            return self._evaluate_program(meanfunc, covfunc, likfunc, x, y, m, sn2, nargout)
        kind, md, hyp = spec
        eng = self._get_sharded() if self._wants_sharding(n) else self._get_engine()
        eng.set_data(x)
        nlZ, alpha, dcov, dlik = eng.exact_eval(kind, md, hyp, likfunc.hyp[0], y - m, nargout > 2)
        post = postStruct()
        post.alpha = alpha
        post.sW = np.ones((n, 1)) / np.sqrt(sn2)
        post._L = None
        post._engine, post._epoch, post._n = eng, eng.epoch, n
        post._spec = ('exact', kind, md, tuple(hyp), float(likfunc.hyp[0]))
        _seal(post, x)
        return self._pack(post, nlZ, dcov, dlik, meanfunc, covfunc, likfunc, x, alpha, nargout)

    @staticmethod
    def _pack(post, nlZ, dcov, dlik, meanfunc, covfunc, likfunc, x, alpha, nargout):
        if nargout > 1:
            if nargout > 2:
                dnlZ = dnlZStruct(meanfunc, covfunc, likfunc)
                dnlZ.lik = [np.float64(dlik[0])]
                dnlZ.cov = [np.float64(v) for v in dcov]
                _mean_derivs(meanfunc, x, alpha, dnlZ)
                return post, nlZ, dnlZ
            return post, nlZ
        return post

    def _evaluate_program(self, meanfunc, covfunc, likfunc, x, y, m, sn2, nargout):
        """Composite kernels and the kernels without a dedicated build: the covariance PROGRAM (cov._device_prog) is
        evaluated on the device inside the same single foreign call - matrix build, factorisation, solves, and one
        fused pass for ALL hyper-parameter derivatives of all components (csrc/covprog.cu).  A cov.Pre leaf reads its
        training matrix from device memory (uploaded here, gpk_set_pre)."""
        n = x.shape[0]
        prog = covfunc._device_prog()
        pres = covfunc._pre_leaves()
        if prog is None or len(pres) > 1:
            raise Exception('%s: this covariance function has no device implementation (there is no CPU fallback)'
                            % type(covfunc).__name__)
        nodes, hyp = prog
        eng = self._get_engine()
        eng.set_data(x)
        if pres:
            M2 = np.asarray(pres[0].M2, dtype=np.float64)
            if M2.shape != (n, n):
                raise Exception('cov.Pre: the training matrix M2 must be (n,n) for n training inputs')
            eng.set_pre(M2)
        nlZ, alpha, dcov, dlik = eng.exact_eval_prog(nodes, hyp, likfunc.hyp[0], y - m, nargout > 2)
        post = postStruct()
        post.alpha = alpha
        post.sW = np.ones((n, 1)) / np.sqrt(sn2)
        post._L = None
        post._engine, post._epoch, post._n = eng, eng.epoch, n
        # a posterior with a Pre leaf cannot be predicted from on the device (its cross-covariances are host matrices)
        post._spec = None if pres else ('exact', 'prog', tuple(nodes), tuple(hyp), float(likfunc.hyp[0]))
        _seal(post, x)
        if pres:
            post._x = None                       # no rebuild path either: fetch the factor while it is resident
            post.L
        return self._pack(post, nlZ, dcov, dlik, meanfunc, covfunc, likfunc, x, alpha, nargout)


class EP(Inference):
    """Expectation Propagation for binary classification with lik.Erf (Core/inf.py:723-806).

    The whole evaluation (kernel matrix, sequential site updates with rank-1 posterior updates, the
    per-sweep refactorisation, posterior parameters and derivatives) is one call, gpk_ep_eval.  Like the
    reference, the site parameters of the previous call are tried as the starting point (last_ttau/last_tnu)."""

    def __init__(self):
        self.name = 'Expectation Propagation'
        self.last_ttau = None
        self.last_tnu = None
        self.last_sweeps = 0
        self._engine = None

    def evaluate(self, meanfunc, covfunc, likfunc, x, y, nargout=1):
        if not isinstance(likfunc, lik.Erf):
            raise Exception('EP on the GPU path supports the Erf likelihood (binary classification) only')
        spec = covfunc._device_spec() if not isinstance(covfunc, cov.FITCOfKernel) else None
        if spec is None:
            raise Exception('EP on the GPU path needs cov.RBF, cov.RBFard or cov.Matern')
        kind, md, hyp = spec
        n = x.shape[0]
        m = meanfunc.getMean(x)
        eng = self._get_engine()
        eng.set_data(x)
        warm = self.last_ttau is not None and len(self.last_ttau) == n
        nlZ, alpha, sW, dcov, dlz, ttau, tnu, sweeps = eng.ep_eval(
            kind, md, hyp, m, y, self.last_ttau if warm else None, self.last_tnu if warm else None, warm, nargout > 2)
        if sweeps == 10:
            logging.getLogger(__name__).warning("maximum number of sweeps reached in function infEP")
        self.last_ttau, self.last_tnu, self.last_sweeps = ttau, tnu, sweeps
        post = postStruct()
        post.alpha = alpha
        post.sW = sW
        post._L = None
        post._engine, post._epoch, post._n = eng, eng.epoch, n
        post._spec = ('ep', kind, md, tuple(hyp), 0.0)
        _seal(post, x)
        post.L                                    # EP evaluations take seconds: fetch the factor now (no rebuild path)
        if nargout > 1:
            if nargout > 2:
                dnlZ = dnlZStruct(meanfunc, covfunc, likfunc)
                dnlZ.cov = [np.float64(v) for v in dcov]
                dnlZ.lik = []
                for i in range(len(meanfunc.hyp)):
                    dnlZ.mean[i] = np.float64(-np.dot(dlz.T, meanfunc.getDerMatrix(x, i))[0, 0])
                return post, nlZ, dnlZ
            return post, nlZ
        return post


class FITC_Exact(Inference):
    """FITC approximation with Gaussian likelihood (Core/inf.py:387-455)."""

    def __init__(self):
        self.name = 'FICT exact inference'
        self._engine = None

    def evaluate(self, meanfunc, covfunc, likfunc, x, y, nargout=1):
        if not isinstance(likfunc, lik.Gauss):
            raise Exception('Exact inference only possible with Gaussian likelihood')
        if not isinstance(covfunc, cov.FITCOfKernel):
            raise Exception('Only covFITC supported.')
        spec = covfunc._device_spec()
        if spec is None:
            raise Exception('FITC on the GPU path needs cov.RBF, cov.RBFard or cov.Matern')
        kind, md, hyp = spec
        xu = covfunc.inducingInput
        if xu.shape[1] != x.shape[1]:
            raise Exception('Dimensionality of inducing inputs must match training inputs')
        n, D = x.shape
        m = meanfunc.getMean(x)
        sn2 = np.exp(2 * likfunc.hyp[0])
        eng = self._get_engine()
        eng.set_data(x)
        nlZ, alpha, Lp, dcov, dlik, al = eng.fitc_eval(kind, md, hyp, likfunc.hyp[0], xu, y - m, nargout > 2)
        post = postStruct()
        post.alpha = alpha
        post.sW = np.ones((n, 1)) / np.sqrt(sn2)
        post.L = Lp
        post._engine, post._epoch, post._n = eng, eng.epoch, xu.shape[0]
        post._spec = ('fitc', kind, md, tuple(hyp), float(likfunc.hyp[0]))
        _seal(post, x)
        if nargout > 1:
            if nargout > 2:
                dnlZ = dnlZStruct(meanfunc, covfunc, likfunc)
                dnlZ.cov = [np.float64(v) for v in dcov]
                dnlZ.lik = [np.float64(dlik[0])]
                _mean_derivs(meanfunc, x, al, dnlZ)
                return post, nlZ, dnlZ
            return post, nlZ
        return post
